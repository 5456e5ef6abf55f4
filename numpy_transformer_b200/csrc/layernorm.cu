// LayerNorm forward/backward (net/layernorm.py:88-135): one warp per row, shuffle reductions.
#include "common.cuh"

namespace nt {

// y = (x-mean)*rstd*gamma + beta (+ post_add[row % period]); mean/rstd saved for backward.
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int M, int D, float eps, const float* __restrict__ post_add,
                                                     int period) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (long long)warp * D;
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += xr[c];
  const float mean = warp_sum(s) / D;
  float v = 0.f;
  for (int c = lane; c < D; c += 32) { float t = xr[c] - mean; v += t * t; }
  const float var = warp_sum(v) / D;                 // biased, np.var
  const float rstd = 1.0f / sqrtf(var + eps);
  if (lane == 0) { mean_out[warp] = mean; rstd_out[warp] = rstd; }
  float* yr = y + (long long)warp * D;
  const float* pa = post_add ? post_add + (long long)(warp % period) * D : nullptr;
  for (int c = lane; c < D; c += 32) {
    float o = (xr[c] - mean) * rstd * gamma[c] + beta[c];
    if (pa) o += pa[c];
    yr[c] = o;
  }
}

// dx = rstd*(g*dy - mean(g*dy) - xhat*mean(g*dy*xhat)) (+ addend)
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ gamma, float* __restrict__ dx,
                                                        int M, int D, const float* __restrict__ addend) {
  pdl_prologue();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const long long off = (long long)warp * D;
  const float mu = mean[warp], rs = rstd[warp];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    float gdy = gamma[c] * dy[off + c];
    float xh = (x[off + c] - mu) * rs;
    s1 += gdy;
    s2 += gdy * xh;
  }
  s1 = warp_sum(s1) / D;
  s2 = warp_sum(s2) / D;
  for (int c = lane; c < D; c += 32) {
    float gdy = gamma[c] * dy[off + c];
    float xh = (x[off + c] - mu) * rs;
    float o = rs * (gdy - s1 - xh * s2);
    if (addend) o += addend[off + c];
    dx[off + c] = o;
  }
}

// dgamma[c] += sum_r xhat*dy ; dbeta[c] += sum_r dy   (column reductions, two-stage: block partials + atomics)
__global__ void ln_bwd_param_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                    const float* __restrict__ mean, const float* __restrict__ rstd,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D,
                                    int rows_per_block) {
  pdl_prologue();
  __shared__ float rg[8][33], rb[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float sg = 0.f, sb = 0.f;
  if (c < D)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) {
      float d = dy[(long long)r * D + c];
      float xh = (x[(long long)r * D + c] - mean[r]) * rstd[r];
      sg += xh * d;
      sb += d;
    }
  rg[threadIdx.y][threadIdx.x] = sg;
  rb[threadIdx.y][threadIdx.x] = sb;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float tg = 0.f, tb = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) { tg += rg[i][threadIdx.x]; tb += rb[i][threadIdx.x]; }
    atomicAdd(dgamma + c, tg);
    atomicAdd(dbeta + c, tb);
  }
}

// Fused backward: one pass over x and dy produces dx AND the dgamma/dbeta partials.  A CTA of 8 warps owns
// LN_ROWS rows; each lane keeps its columns (c = lane + 32k) of x/dy in registers, so the row reductions are
// warp shuffles and the column reductions are register accumulators -> smem across warps -> one atomicAdd
// per column per CTA.  NC = columns per lane (D <= 32*NC).
constexpr int LN_ROWS = 32;
template <int NC>
__global__ void __launch_bounds__(256) ln_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, float* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                                                           int D, const float* __restrict__ addend) {
  pdl_prologue();
  extern __shared__ float red[];                 // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * LN_ROWS;
  float g[NC], ag[NC], ab[NC];
#pragma unroll
  for (int k = 0; k < NC; k++) {
    const int c = lane + 32 * k;
    g[k] = c < D ? gamma[c] : 0.f;
    ag[k] = 0.f;
    ab[k] = 0.f;
  }
  const float invD = 1.0f / D;
  for (int r = r0 + warp; r < min(M, r0 + LN_ROWS); r += 8) {
    const long long off = (long long)r * D;
    const float mu = mean[r], rs = rstd[r];
    float dyv[NC], xh[NC];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NC; k++) {
      const int c = lane + 32 * k;
      dyv[k] = c < D ? dy[off + c] : 0.f;
      xh[k] = c < D ? (x[off + c] - mu) * rs : 0.f;
      const float gdy = g[k] * dyv[k];
      s1 += gdy;
      s2 += gdy * xh[k];
      ag[k] = fmaf(xh[k], dyv[k], ag[k]);
      ab[k] += dyv[k];
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    if (dx) {
#pragma unroll
      for (int k = 0; k < NC; k++) {
        const int c = lane + 32 * k;
        if (c < D) {
          float o = rs * (g[k] * dyv[k] - s1 - xh[k] * s2);
          if (addend) o += addend[off + c];
          dx[off + c] = o;
        }
      }
    }
  }
  if (dgamma) {
#pragma unroll
    for (int k = 0; k < NC; k++) {
      const int c = lane + 32 * k;
      if (c < D) { red[(warp * 2) * D + c] = ag[k]; red[(warp * 2 + 1) * D + c] = ab[k]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += 256) {
      float tg = 0.f, tb = 0.f;
#pragma unroll
      for (int w = 0; w < 8; w++) { tg += red[(w * 2) * D + c]; tb += red[(w * 2 + 1) * D + c]; }
      atomicAdd(dgamma + c, tg);
      atomicAdd(dbeta + c, tb);
    }
  }
}

template <int NC>
static int launch_ln_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                               float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend) {
  const size_t smem = sizeof(float) * 16 * (size_t)D;
  NT_CHECK(nt::launch_k(ln_bwd_fused_kernel<NC>, dim3(ceil_div(M, LN_ROWS)), dim3(256), smem, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D,
                                                                         addend));
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
using namespace nt;

extern "C" {

int nt_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                     int M, int D, float eps, const float* post_add, int period) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && D > 0, "nt_layernorm_fwd: bad shape M=%d D=%d", M, D);
  NT_REQUIRE(!post_add || period > 0, "nt_layernorm_fwd: post_add needs period > 0");
  if (M == 0) return 0;
  NT_CHECK(nt::launch_k(ln_fwd_kernel, dim3(ceil_div((long long)M * 32, 256)), dim3(256), 0, x, gamma, beta, y, mean, rstd, M, D, eps,
                                                                        post_add, period));
  NT_LAUNCH_OK();
  return 0;
}

int nt_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && D > 0, "nt_layernorm_bwd: bad shape M=%d D=%d", M, D);
  if (M == 0) return 0;
  if (D <= 768 && (dgamma != nullptr) == (dbeta != nullptr)) {
    if (D <= 128) return launch_ln_bwd_fused<4>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend);
    if (D <= 256) return launch_ln_bwd_fused<8>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend);
    if (D <= 384) return launch_ln_bwd_fused<12>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend);
    return launch_ln_bwd_fused<24>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend);
  }
  if (dgamma && dbeta) {
    int rows_per_block = M >= 8192 ? 512 : 128;
    dim3 grid(ceil_div(D, 32), ceil_div(M, rows_per_block));
    NT_CHECK(nt::launch_k(ln_bwd_param_kernel, dim3(grid), dim3(dim3(32, 8)), 0, dy, x, mean, rstd, dgamma, dbeta, M, D, rows_per_block));
    NT_LAUNCH_OK();
  }
  if (dx) {
    NT_CHECK(nt::launch_k(ln_bwd_dx_kernel, dim3(ceil_div((long long)M * 32, 256)), dim3(256), 0, dy, x, mean, rstd, gamma, dx, M, D, addend));
    NT_LAUNCH_OK();
  }
  return 0;
}

}  // extern "C"
