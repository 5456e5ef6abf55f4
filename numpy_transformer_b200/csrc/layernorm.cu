// LayerNorm forward/backward (net/layernorm.py:88-135): one warp per row, shuffle reductions.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace nt {

// y = (x-mean)*rstd*gamma + beta (+ post_add[row % period]); mean/rstd saved for backward.
__global__ void __launch_bounds__(256) ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, float* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     int M, int D, float eps, const float* __restrict__ post_add,
                                                     int period, __nv_bfloat16* __restrict__ y_hi,
                                                     __nv_bfloat16* __restrict__ y_lo) {
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
    if (y_hi) {                                    // operand of the next BF16x3 GEMM: hi = bf16(y), lo = bf16(y - hi)
      const __nv_bfloat16 h = __float2bfloat16_rn(o);
      y_hi[(long long)warp * D + c] = h;
      y_lo[(long long)warp * D + c] = __float2bfloat16_rn(o - __bfloat162float(h));
    }
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
// `rows` rows; each lane keeps NV float4 column groups (c = 4*(lane + 32k)) of x / dy / addend in registers,
// loaded with 128-bit accesses issued together, so the row reductions are warp shuffles and the column
// reductions are register accumulators -> smem across warps -> one atomicAdd per column per CTA.  D % 4 == 0.
template <int NV>
__global__ void __launch_bounds__(256) ln_bwd_fused_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           const float* __restrict__ gamma, float* __restrict__ dx,
                                                           float* __restrict__ dgamma, float* __restrict__ dbeta, int M,
                                                           int D, const float* __restrict__ addend, int LN_ROWS,
                                                           uint2* __restrict__ dx_hi, uint2* __restrict__ dx_lo) {
  pdl_prologue();
  extern __shared__ float red[];                 // [8 warps][2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = blockIdx.x * LN_ROWS;
  const int D4 = D >> 2;
  float4 g[NV], ag[NV], ab[NV];
#pragma unroll
  for (int k = 0; k < NV; k++) {
    const int c4 = lane + 32 * k;
    g[k] = c4 < D4 ? reinterpret_cast<const float4*>(gamma)[c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    ag[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    ab[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invD = 1.0f / D;
  const int rend = min(M, r0 + LN_ROWS);
  for (int r = r0 + warp; r < rend; r += 8) {
    const long long off4 = (long long)r * D4;
    float4 dyv[NV], xv[NV], av[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c4 = lane + 32 * k;
      const bool ok = c4 < D4;
      dyv[k] = ok ? reinterpret_cast<const float4*>(dy)[off4 + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      xv[k] = ok ? reinterpret_cast<const float4*>(x)[off4 + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
      av[k] = (ok && addend) ? reinterpret_cast<const float4*>(addend)[off4 + c4] : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const bool ok = (lane + 32 * k) < D4;
      float4& xh = xv[k];
      xh.x = ok ? (xh.x - mu) * rs : 0.f; xh.y = ok ? (xh.y - mu) * rs : 0.f;
      xh.z = ok ? (xh.z - mu) * rs : 0.f; xh.w = ok ? (xh.w - mu) * rs : 0.f;
      const float4 d = dyv[k];
      const float gx = g[k].x * d.x, gy = g[k].y * d.y, gz = g[k].z * d.z, gw = g[k].w * d.w;
      s1 += gx + gy + gz + gw;
      s2 += gx * xh.x + gy * xh.y + gz * xh.z + gw * xh.w;
      ag[k].x = fmaf(xh.x, d.x, ag[k].x); ag[k].y = fmaf(xh.y, d.y, ag[k].y);
      ag[k].z = fmaf(xh.z, d.z, ag[k].z); ag[k].w = fmaf(xh.w, d.w, ag[k].w);
      ab[k].x += d.x; ab[k].y += d.y; ab[k].z += d.z; ab[k].w += d.w;
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    if (dx) {
#pragma unroll
      for (int k = 0; k < NV; k++) {
        const int c4 = lane + 32 * k;
        if (c4 < D4) {
          const float4 d = dyv[k], xh = xv[k];
          float4 o;
          o.x = rs * (g[k].x * d.x - s1 - xh.x * s2) + av[k].x;
          o.y = rs * (g[k].y * d.y - s1 - xh.y * s2) + av[k].y;
          o.z = rs * (g[k].z * d.z - s1 - xh.z * s2) + av[k].z;
          o.w = rs * (g[k].w * d.w - s1 - xh.w * s2) + av[k].w;
          reinterpret_cast<float4*>(dx)[off4 + c4] = o;
          if (dx_hi) {
            __nv_bfloat162 h0 = __floats2bfloat162_rn(o.x, o.y), h1 = __floats2bfloat162_rn(o.z, o.w);
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            __nv_bfloat162 l0 = __floats2bfloat162_rn(o.x - f0.x, o.y - f0.y), l1 = __floats2bfloat162_rn(o.z - f1.x, o.w - f1.y);
            dx_hi[off4 + c4] = make_uint2(*reinterpret_cast<uint32_t*>(&h0), *reinterpret_cast<uint32_t*>(&h1));
            dx_lo[off4 + c4] = make_uint2(*reinterpret_cast<uint32_t*>(&l0), *reinterpret_cast<uint32_t*>(&l1));
          }
        }
      }
    }
  }
  if (dgamma) {
#pragma unroll
    for (int k = 0; k < NV; k++) {
      const int c4 = lane + 32 * k;
      if (c4 < D4) {
        reinterpret_cast<float4*>(red + (warp * 2) * D)[c4] = ag[k];
        reinterpret_cast<float4*>(red + (warp * 2 + 1) * D)[c4] = ab[k];
      }
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

template <int NV>
static int launch_ln_bwd_fused(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                               float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend, void* dx_hi,
                               void* dx_lo) {
  const size_t smem = sizeof(float) * 16 * (size_t)D;
  // rows per CTA: every CTA ends with 2*D same-address atomics, so large M uses taller CTAs
  int rows = 32;
  if (const char* e = getenv("NTB200_LN_ROWS")) rows = atoi(e);
  else if (M >= 32768) rows = 128;
  NT_CHECK(nt::launch_k(ln_bwd_fused_kernel<NV>, dim3(ceil_div(M, rows)), dim3(256), smem, dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D,
                                                                         addend, rows, (uint2*)dx_hi, (uint2*)dx_lo));
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
using namespace nt;

extern "C" {

static int ln_fwd_impl(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M, int D,
                       float eps, const float* post_add, int period, void* y_hi, void* y_lo) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && D > 0, "nt_layernorm_fwd: bad shape M=%d D=%d", M, D);
  NT_REQUIRE(!post_add || period > 0, "nt_layernorm_fwd: post_add needs period > 0");
  NT_REQUIRE((y_hi != nullptr) == (y_lo != nullptr), "nt_layernorm_fwd_split: pass both split planes or neither");
  if (M == 0) return 0;
  NT_CHECK(nt::launch_k(ln_fwd_kernel, dim3(ceil_div((long long)M * 32, 256)), dim3(256), 0, x, gamma, beta, y, mean, rstd, M, D, eps,
                                                                        post_add, period, (__nv_bfloat16*)y_hi, (__nv_bfloat16*)y_lo));
  NT_LAUNCH_OK();
  return 0;
}

int nt_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                     int M, int D, float eps, const float* post_add, int period) {
  NT_ENTRY();
  NT_FUSE_TRY(nt::fuse_ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, post_add));
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, M, D, eps, post_add, period, nullptr, nullptr);
}

int nt_layernorm_fwd_split(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                           int M, int D, float eps, const float* post_add, int period, void* y_hi, void* y_lo) {
  return ln_fwd_impl(x, gamma, beta, y, mean, rstd, M, D, eps, post_add, period, y_hi, y_lo);
}

static int ln_bwd_impl(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                       float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend, void* dx_hi, void* dx_lo);

int nt_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                     float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend) {
  NT_ENTRY();
  NT_FUSE_TRY(nt::fuse_ln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend));
  return ln_bwd_impl(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, nullptr, nullptr);
}

int nt_layernorm_bwd_split(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                           float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend, void* dx_hi, void* dx_lo) {
  return ln_bwd_impl(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, dx_hi, dx_lo);
}

static int ln_bwd_impl(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                       float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend, void* dx_hi, void* dx_lo) {
  NT_ENTRY();
  NT_REQUIRE((dx_hi != nullptr) == (dx_lo != nullptr), "nt_layernorm_bwd_split: pass both split planes or neither");
  NT_REQUIRE(M >= 0 && D > 0, "nt_layernorm_bwd: bad shape M=%d D=%d", M, D);
  if (M == 0) return 0;
  static int split_big = -1;
  if (split_big < 0) { const char* e = getenv("NTB200_LN_SPLIT"); split_big = e ? atoi(e) : 0; }
  const bool vec_ok = (D % 4 == 0) && ((((uintptr_t)dy | (uintptr_t)x | (uintptr_t)dx | (uintptr_t)addend | (uintptr_t)gamma) & 15) == 0);
  if (D <= 768 && vec_ok && (dgamma != nullptr) == (dbeta != nullptr) && !(split_big && M >= 32768)) {
    if (D <= 128) return launch_ln_bwd_fused<1>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, dx_hi, dx_lo);
    if (D <= 256) return launch_ln_bwd_fused<2>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, dx_hi, dx_lo);
    if (D <= 384) return launch_ln_bwd_fused<3>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, dx_hi, dx_lo);
    return launch_ln_bwd_fused<6>(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend, dx_hi, dx_lo);
  }
  NT_REQUIRE(!dx_hi, "nt_layernorm_bwd_split: split output needs the fused kernel (D %% 4 == 0, D <= 768, aligned pointers)");
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
