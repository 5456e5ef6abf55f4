// Softmax rows and the attention core (attention.py:31-46,63-84; gpt/attdecoderblock.py:52-69,79-99).
// Two paths: one fused CTA per (sample, head) with everything in shared memory when N<=64 and
// dh<=64 (the ViT shapes), otherwise batched GEMMs + masked row softmax (the T=256 GPT shapes).
#include "common.cuh"
#include <cmath>

namespace nt {

// p = softmax(x + mask) along rows; one warp per row. mask_kind as in ntb200.h; qi = row % N.
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ x, float* __restrict__ p,
                                                           long long rows, int cols, const float* __restrict__ mask,
                                                           int mask_kind, int N) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * cols;
  float* pr = p + row * cols;
  const int qi = (mask_kind != NT_MASK_NONE) ? (int)(row % N) : 0;
  auto val = [&](int j) -> float {
    float v = xr[j];
    if (mask_kind == NT_MASK_CAUSAL) { if (j > qi) v = -INFINITY; }
    else if (mask_kind == NT_MASK_DENSE) v += mask[(long long)qi * cols + j];
    return v;
  };
  float mx = -INFINITY;
  for (int j = lane; j < cols; j += 32) mx = fmaxf(mx, val(j));
  mx = warp_max(mx);
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s += expf(val(j) - mx);
  s = warp_sum(s);
  const float inv = 1.0f / s;
  for (int j = lane; j < cols; j += 32) pr[j] = expf(val(j) - mx) * inv;
}

// dx = p * (dp - sum(dp*p))  (closed form of the per-row Jacobian product, activation.py:86-91)
__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const float* dp, const float* __restrict__ p,
                                                               float* dx, long long rows, int cols) {
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* dr = dp + row * cols;
  const float* pr = p + row * cols;
  float s = 0.f;
  for (int j = lane; j < cols; j += 32) s += dr[j] * pr[j];
  s = warp_sum(s);
  for (int j = lane; j < cols; j += 32) dx[row * cols + j] = pr[j] * (dr[j] - s);
}

// ---------------------------------------------------------------- fused small-N kernels
__global__ void __launch_bounds__(128) attn_fwd_small_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                             int mask_kind, float* __restrict__ out,
                                                             float* __restrict__ P, float* __restrict__ S, int N, int H,
                                                             int dh, const float* __restrict__ residual) {
  extern __shared__ float sm[];
  const int dhp = dh + 1, Np = N + 1;
  float* sQ = sm;
  float* sK = sQ + N * dh;
  float* sV = sK + N * dhp;
  float* sP = sV + N * dh;
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * dh;
  const int tid = threadIdx.x, nt_ = blockDim.x;
  const float* base = qkv + (long long)b * N * 3 * D + h * dh;
  for (int e = tid; e < N * dh; e += nt_) {
    const int i = e / dh, d = e % dh;
    const float* r = base + (long long)i * 3 * D + d;
    sQ[i * dh + d] = r[0];
    sK[i * dhp + d] = r[D];
    sV[i * dh + d] = r[2 * D];
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  for (int e = tid; e < N * N; e += nt_) {
    const int i = e / N, j = e % N;
    float acc = 0.f;
    for (int d = 0; d < dh; d++) acc = fmaf(sQ[i * dh + d], sK[j * dhp + d], acc);
    acc *= scale;                                           // attention.py:37
    if (S) S[(long long)bh * N * N + e] = acc;              // att_col, attdecoderblock.py:58-59
    if (mask_kind == NT_MASK_CAUSAL) { if (j > i) acc = -INFINITY; }
    else if (mask_kind == NT_MASK_DENSE) acc += mask[i * N + j];   // attention.py:38-39
    sP[i * Np + j] = acc;
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = nt_ >> 5;
  for (int i = warp; i < N; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sP[i * Np + j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) { float e_ = expf(sP[i * Np + j] - mx); sP[i * Np + j] = e_; s += e_; }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < N; j += 32) {
      float pv = sP[i * Np + j] * inv;
      sP[i * Np + j] = pv;
      P[((long long)bh * N + i) * N + j] = pv;              // atg__
    }
  }
  __syncthreads();
  for (int e = tid; e < N * dh; e += nt_) {
    const int i = e / dh, d = e % dh;
    float acc = 0.f;
    for (int j = 0; j < N; j++) acc = fmaf(sP[i * Np + j], sV[j * dh + d], acc);
    const long long o = ((long long)b * N + i) * D + h * dh + d;
    out[o] = residual ? acc + residual[o] : acc;
  }
}

__global__ void __launch_bounds__(128) attn_bwd_small_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                             const float* __restrict__ P, float* __restrict__ dqkv,
                                                             int N, int H, int dh) {
  extern __shared__ float sm[];
  const int dhp = dh + 1, Np = N + 1;
  float* sQ = sm;                  // [N][dh]
  float* sK = sQ + N * dh;         // [N][dh]
  float* sV = sK + N * dh;         // [N][dh+1]
  float* sdO = sV + N * dhp;       // [N][dh]
  float* sP = sdO + N * dh;        // [N][N+1]
  float* sdS = sP + N * Np;        // [N][N+1]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * dh;
  const int tid = threadIdx.x, nt_ = blockDim.x;
  const float* base = qkv + (long long)b * N * 3 * D + h * dh;
  for (int e = tid; e < N * dh; e += nt_) {
    const int i = e / dh, d = e % dh;
    const float* r = base + (long long)i * 3 * D + d;
    sQ[i * dh + d] = r[0];
    sK[i * dh + d] = r[D];
    sV[i * dhp + d] = r[2 * D];
    sdO[i * dh + d] = dout[((long long)b * N + i) * D + h * dh + d];
  }
  for (int e = tid; e < N * N; e += nt_) sP[(e / N) * Np + e % N] = P[(long long)bh * N * N + e];
  __syncthreads();
  // dP = dO V^T   (attention.py:74)
  for (int e = tid; e < N * N; e += nt_) {
    const int i = e / N, j = e % N;
    float acc = 0.f;
    for (int d = 0; d < dh; d++) acc = fmaf(sdO[i * dh + d], sV[j * dhp + d], acc);
    sdS[i * Np + j] = acc;
  }
  __syncthreads();
  // dS = P o (dP - rowsum(dP o P))   (attention.py:75)
  const int warp = tid >> 5, lane = tid & 31, nw = nt_ >> 5;
  for (int i = warp; i < N; i += nw) {
    float s = 0.f;
    for (int j = lane; j < N; j += 32) s += sdS[i * Np + j] * sP[i * Np + j];
    s = warp_sum(s);
    for (int j = lane; j < N; j += 32) sdS[i * Np + j] = sP[i * Np + j] * (sdS[i * Np + j] - s);
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  float* obase = dqkv + (long long)b * N * 3 * D + h * dh;
  for (int e = tid; e < N * dh; e += nt_) {
    const int r = e / dh, d = e % dh;
    float aq = 0.f, ak = 0.f, av = 0.f;
    for (int t = 0; t < N; t++) {
      aq = fmaf(sdS[r * Np + t], sK[t * dh + d], aq);       // dQ = dS K       (attention.py:78)
      ak = fmaf(sdS[t * Np + r], sQ[t * dh + d], ak);       // dK = dS^T Q     (attention.py:79)
      av = fmaf(sP[t * Np + r], sdO[t * dh + d], av);       // dV = P^T dO     (attention.py:76)
    }
    float* o = obase + (long long)r * 3 * D + d;
    o[0] = aq * scale;
    o[D] = ak * scale;
    o[2 * D] = av;
  }
}

static size_t fwd_small_smem(int N, int dh) { return sizeof(float) * ((size_t)N * dh * 2 + (size_t)N * (dh + 1) + (size_t)N * (N + 1)); }
static size_t bwd_small_smem(int N, int dh) { return sizeof(float) * ((size_t)N * dh * 3 + (size_t)N * (dh + 1) + 2 * (size_t)N * (N + 1)); }

static int softmax_rows(const float* x, float* p, long long rows, int cols, const float* mask, int mask_kind, int N) {
  if (rows == 0) return 0;
  softmax_rows_kernel<<<ceil_div(rows * 32, 256), 256, 0, g_stream>>>(x, p, rows, cols, mask, mask_kind, N);
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
using namespace nt;

extern "C" {

int nt_softmax_fwd(const float* x, float* p, int rows, int cols) {
  NT_ENTRY();
  NT_REQUIRE(rows >= 0 && cols > 0, "nt_softmax_fwd: bad shape");
  return softmax_rows(x, p, rows, cols, nullptr, NT_MASK_NONE, 1);
}

int nt_softmax_bwd(const float* dp, const float* p, float* dx, int rows, int cols) {
  NT_ENTRY();
  NT_REQUIRE(rows >= 0 && cols > 0, "nt_softmax_bwd: bad shape");
  if (rows == 0) return 0;
  softmax_bwd_rows_kernel<<<ceil_div((long long)rows * 32, 256), 256, 0, g_stream>>>(dp, p, dx, rows, cols);
  NT_LAUNCH_OK();
  return 0;
}

int nt_attention_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N,
                     int H, int dh, const float* residual) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0, "nt_attention_fwd: bad shape");
  NT_REQUIRE(mask_kind != NT_MASK_DENSE || mask, "nt_attention_fwd: dense mask_kind needs a mask");
  if (B == 0) return 0;
  const int D = H * dh;
  if (N <= 64 && dh <= 64) {
    size_t smem = fwd_small_smem(N, dh);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      NT_CHECK(cudaFuncSetAttribute(attn_fwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    attn_fwd_small_kernel<<<B * H, 128, smem, g_stream>>>(qkv, mask, mask_kind, out, P, S, N, H, dh, residual);
    NT_LAUNCH_OK();
    return 0;
  }
  // general path: S = scale * Q K^T (batched over (b,h)) -> masked softmax -> O = P V
  GemmDesc g{};
  g.A = qkv; g.B = qkv + D; g.C = S ? S : P;
  g.M = N; g.N = N; g.K = dh;
  g.sAm = 3 * D; g.sAk = 1; g.sBk = 1; g.sBn = 3 * D; g.ldc = N;
  g.nb0 = B; g.nb1 = H;
  g.bA0 = (long long)N * 3 * D; g.bA1 = dh; g.bB0 = g.bA0; g.bB1 = dh;
  g.bC0 = (long long)H * N * N; g.bC1 = (long long)N * N;
  g.ep.alpha = 1.0f / sqrtf((float)dh);
  int rc = gemm_simt(g);
  if (rc) return rc;
  rc = softmax_rows(S ? S : P, P, (long long)B * H * N, N, mask, mask_kind, N);
  if (rc) return rc;
  GemmDesc o{};
  o.A = P; o.B = qkv + 2 * D; o.C = out;
  o.M = N; o.N = dh; o.K = N;
  o.sAm = N; o.sAk = 1; o.sBk = 3 * D; o.sBn = 1; o.ldc = D;
  o.nb0 = B; o.nb1 = H;
  o.bA0 = (long long)H * N * N; o.bA1 = (long long)N * N;
  o.bB0 = (long long)N * 3 * D; o.bB1 = dh;
  o.bC0 = (long long)N * D; o.bC1 = dh;
  o.ep.addend = residual;
  return gemm_simt(o);
}

int nt_attention_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, float* scratch, int B, int N,
                     int H, int dh) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0, "nt_attention_bwd: bad shape");
  if (B == 0) return 0;
  const int D = H * dh;
  if (N <= 64 && dh <= 64) {
    size_t smem = bwd_small_smem(N, dh);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
      NT_CHECK(cudaFuncSetAttribute(attn_bwd_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      configured = smem;
    }
    attn_bwd_small_kernel<<<B * H, 128, smem, g_stream>>>(dout, qkv, P, dqkv, N, H, dh);
    NT_LAUNCH_OK();
    return 0;
  }
  NT_REQUIRE(scratch, "nt_attention_bwd: general path needs a B*H*N*N scratch buffer");
  const long long bhs = (long long)N * N;
  const float scale = 1.0f / sqrtf((float)dh);
  // dP = dO V^T
  GemmDesc g{};
  g.A = dout; g.B = qkv + 2 * D; g.C = scratch;
  g.M = N; g.N = N; g.K = dh;
  g.sAm = D; g.sAk = 1; g.sBk = 1; g.sBn = 3 * D; g.ldc = N;
  g.nb0 = B; g.nb1 = H;
  g.bA0 = (long long)N * D; g.bA1 = dh; g.bB0 = (long long)N * 3 * D; g.bB1 = dh;
  g.bC0 = H * bhs; g.bC1 = bhs;
  int rc = gemm_simt(g);
  if (rc) return rc;
  // dS in place
  softmax_bwd_rows_kernel<<<ceil_div((long long)B * H * N * 32, 256), 256, 0, g_stream>>>(scratch, P, scratch,
                                                                                         (long long)B * H * N, N);
  NT_LAUNCH_OK();
  // dV = P^T dO
  GemmDesc v{};
  v.A = P; v.B = dout; v.C = dqkv + 2 * D;
  v.M = N; v.N = dh; v.K = N;
  v.sAm = 1; v.sAk = N; v.sBk = D; v.sBn = 1; v.ldc = 3 * D;
  v.nb0 = B; v.nb1 = H;
  v.bA0 = H * bhs; v.bA1 = bhs; v.bB0 = (long long)N * D; v.bB1 = dh;
  v.bC0 = (long long)N * 3 * D; v.bC1 = dh;
  rc = gemm_simt(v);
  if (rc) return rc;
  // dQ = scale * dS K
  GemmDesc q = v;
  q.A = scratch; q.B = qkv + D; q.C = dqkv;
  q.sAm = N; q.sAk = 1; q.sBk = 3 * D; q.sBn = 1;
  q.bB0 = (long long)N * 3 * D; q.bB1 = dh;
  q.ep.alpha = scale;
  rc = gemm_simt(q);
  if (rc) return rc;
  // dK = scale * dS^T Q
  GemmDesc k = q;
  k.B = qkv; k.C = dqkv + D;
  k.sAm = 1; k.sAk = N;
  return gemm_simt(k);
}

}  // extern "C"
