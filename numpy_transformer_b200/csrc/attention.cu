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
  pdl_prologue();
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
  pdl_prologue();
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
// One CTA per (sample, head); Q/K/V (and dO, P, dS in backward) live in shared memory.  Every thread owns
// a 4x4 register tile fed by 128-bit LDS (16 FMA per two loads: a 1x4 tile is 5x shared-memory-bandwidth
// bound).  Operands whose tile runs along the token axis are staged transposed ([dh][Np]).  dh % 4 == 0.
constexpr int ATT_THREADS = 256;

__device__ __forceinline__ int pad4(int n) { return (n + 3) & ~3; }

#define NT_FMA4x4(acc, a4, b4)                                                                                   \
  acc[0][0] = fmaf(a4.x, b4.x, acc[0][0]); acc[0][1] = fmaf(a4.x, b4.y, acc[0][1]);                                \
  acc[0][2] = fmaf(a4.x, b4.z, acc[0][2]); acc[0][3] = fmaf(a4.x, b4.w, acc[0][3]);                                \
  acc[1][0] = fmaf(a4.y, b4.x, acc[1][0]); acc[1][1] = fmaf(a4.y, b4.y, acc[1][1]);                                \
  acc[1][2] = fmaf(a4.y, b4.z, acc[1][2]); acc[1][3] = fmaf(a4.y, b4.w, acc[1][3]);                                \
  acc[2][0] = fmaf(a4.z, b4.x, acc[2][0]); acc[2][1] = fmaf(a4.z, b4.y, acc[2][1]);                                \
  acc[2][2] = fmaf(a4.z, b4.z, acc[2][2]); acc[2][3] = fmaf(a4.z, b4.w, acc[2][3]);                                \
  acc[3][0] = fmaf(a4.w, b4.x, acc[3][0]); acc[3][1] = fmaf(a4.w, b4.y, acc[3][1]);                                \
  acc[3][2] = fmaf(a4.w, b4.z, acc[3][2]); acc[3][3] = fmaf(a4.w, b4.w, acc[3][3]);

// NC / DHC: compile-time token count / head width (0 = use the runtime arguments).  With constants the
// index divisions fold to shifts and the contraction loops unroll; the generic instance was
// instruction-bound on integer div/mod (ncu: 26.6 K warp-instructions per CTA for ~7 K of FMA work).
template <int NC, int DHC>
__global__ void __launch_bounds__(ATT_THREADS) attn_fwd_small_kernel(const float* __restrict__ qkv,
                                                                      const float* __restrict__ mask, int mask_kind,
                                                                      float* __restrict__ out, float* __restrict__ P,
                                                                      float* __restrict__ S, int N_rt, int H, int dh_rt,
                                                                      const float* __restrict__ residual) {
  pdl_prologue();
  const int N = NC ? NC : N_rt, dh = DHC ? DHC : dh_rt;
  extern __shared__ __align__(16) float sm[];
  const int N4 = pad4(N), Np = N4 + 4;
  float* sQt = sm;                   // [dh][Np]
  float* sKt = sQt + dh * Np;        // [dh][Np]
  float* sV = sKt + dh * Np;         // [N][dh]
  float* sP = sV + N * dh;           // [N4][Np]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * dh;
  const int tid = threadIdx.x;
  const int dh4 = dh >> 2, ng = N4 >> 2;
  const float* base = qkv + (long long)b * N * 3 * D + h * dh;
  for (int e = tid; e < N4 * dh4; e += ATT_THREADS) {
    const int i = e / dh4, d4 = (e % dh4) * 4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f), k = q;
    if (i < N) {
      const float* r = base + (long long)i * 3 * D + d4;
      q = __ldg(reinterpret_cast<const float4*>(r));
      k = __ldg(reinterpret_cast<const float4*>(r + D));
      *reinterpret_cast<float4*>(sV + i * dh + d4) = __ldg(reinterpret_cast<const float4*>(r + 2 * D));
    }
    sQt[(d4 + 0) * Np + i] = q.x; sQt[(d4 + 1) * Np + i] = q.y; sQt[(d4 + 2) * Np + i] = q.z; sQt[(d4 + 3) * Np + i] = q.w;
    sKt[(d4 + 0) * Np + i] = k.x; sKt[(d4 + 1) * Np + i] = k.y; sKt[(d4 + 2) * Np + i] = k.z; sKt[(d4 + 3) * Np + i] = k.w;
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  for (int e = tid; e < ng * ng; e += ATT_THREADS) {
    const int i0 = (e / ng) * 4, j0 = (e % ng) * 4;
    float acc[4][4] = {};
#pragma unroll 4
    for (int d = 0; d < dh; d++) {
      const float4 q4 = *reinterpret_cast<const float4*>(sQt + d * Np + i0);
      const float4 k4 = *reinterpret_cast<const float4*>(sKt + d * Np + j0);
      NT_FMA4x4(acc, q4, k4)
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int i = i0 + r;
#pragma unroll
      for (int t = 0; t < 4; t++) {
        const int j = j0 + t;
        float v = acc[r][t] * scale;                                               // attention.py:37
        if (i < N && j < N) {
          if (S) S[(long long)bh * N * N + i * N + j] = v;                          // att_col
          if (mask_kind == NT_MASK_CAUSAL) { if (j > i) v = -INFINITY; }
          else if (mask_kind == NT_MASK_DENSE) v += mask[i * N + j];                // attention.py:38-39
        }
        sP[i * Np + j] = v;
      }
    }
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = ATT_THREADS >> 5;
  for (int i = warp; i < N; i += nw) {
    float mx = -INFINITY;
    for (int j = lane; j < N; j += 32) mx = fmaxf(mx, sP[i * Np + j]);
    mx = warp_max(mx);
    float s = 0.f;
    for (int j = lane; j < N; j += 32) { float e_ = expf(sP[i * Np + j] - mx); sP[i * Np + j] = e_; s += e_; }
    s = warp_sum(s);
    const float inv = 1.0f / s;
    for (int j = lane; j < N; j += 32) {
      const float pv = sP[i * Np + j] * inv;
      sP[i * Np + j] = pv;
      P[((long long)bh * N + i) * N + j] = pv;                                     // atg__
    }
  }
  __syncthreads();
  for (int e = tid; e < ng * dh4; e += ATT_THREADS) {
    const int i0 = (e / dh4) * 4, d4 = (e % dh4) * 4;
    float acc[4][4] = {};
    const float* p0 = sP + i0 * Np;
#pragma unroll 2
    for (int j = 0; j < N; j++) {
      const float4 p4 = make_float4(p0[j], p0[Np + j], p0[2 * Np + j], p0[3 * Np + j]);
      const float4 v4 = *reinterpret_cast<const float4*>(sV + j * dh + d4);
      NT_FMA4x4(acc, p4, v4)
    }
#pragma unroll
    for (int r = 0; r < 4; r++) {
      const int i = i0 + r;
      if (i < N) {
        const long long o = ((long long)b * N + i) * D + h * dh + d4;
        float4 v = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        if (residual) {
          const float4 rr = __ldg(reinterpret_cast<const float4*>(residual + o));
          v.x += rr.x; v.y += rr.y; v.z += rr.z; v.w += rr.w;
        }
        *reinterpret_cast<float4*>(out + o) = v;
      }
    }
  }
}

template <int NC, int DHC>
__global__ void __launch_bounds__(ATT_THREADS) attn_bwd_small_kernel(const float* __restrict__ dout,
                                                                      const float* __restrict__ qkv,
                                                                      const float* __restrict__ P,
                                                                      float* __restrict__ dqkv, int N_rt, int H, int dh_rt) {
  pdl_prologue();
  const int N = NC ? NC : N_rt, dh = DHC ? DHC : dh_rt;
  extern __shared__ __align__(16) float sm[];
  const int N4 = pad4(N), Np = N4 + 4;
  float* sQ = sm;                    // [N][dh]
  float* sK = sQ + N * dh;           // [N][dh]
  float* sdO = sK + N * dh;          // [N][dh]
  float* sVt = sdO + N * dh;         // [dh][Np]
  float* sdOt = sVt + dh * Np;       // [dh][Np]
  float* sP = sdOt + dh * Np;        // [N4][Np]
  float* sdS = sP + N4 * Np;         // [N4][Np]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * dh;
  const int tid = threadIdx.x;
  const int dh4 = dh >> 2, ng = N4 >> 2;
  const float* base = qkv + (long long)b * N * 3 * D + h * dh;
  for (int e = tid; e < N4 * dh4; e += ATT_THREADS) {
    const int i = e / dh4, d4 = (e % dh4) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = v;
    if (i < N) {
      const float* r = base + (long long)i * 3 * D + d4;
      *reinterpret_cast<float4*>(sQ + i * dh + d4) = __ldg(reinterpret_cast<const float4*>(r));
      *reinterpret_cast<float4*>(sK + i * dh + d4) = __ldg(reinterpret_cast<const float4*>(r + D));
      v = __ldg(reinterpret_cast<const float4*>(r + 2 * D));
      g = __ldg(reinterpret_cast<const float4*>(dout + ((long long)b * N + i) * D + h * dh + d4));
      *reinterpret_cast<float4*>(sdO + i * dh + d4) = g;
    }
    sVt[(d4 + 0) * Np + i] = v.x; sVt[(d4 + 1) * Np + i] = v.y; sVt[(d4 + 2) * Np + i] = v.z; sVt[(d4 + 3) * Np + i] = v.w;
    sdOt[(d4 + 0) * Np + i] = g.x; sdOt[(d4 + 1) * Np + i] = g.y; sdOt[(d4 + 2) * Np + i] = g.z; sdOt[(d4 + 3) * Np + i] = g.w;
  }
  for (int e = tid; e < N4 * N4; e += ATT_THREADS) {
    const int i = e / N4, j = e % N4;
    sP[i * Np + j] = (i < N && j < N) ? __ldg(P + (long long)bh * N * N + i * N + j) : 0.f;
  }
  __syncthreads();
  // dP = dO V^T   (attention.py:74)
  for (int e = tid; e < ng * ng; e += ATT_THREADS) {
    const int i0 = (e / ng) * 4, j0 = (e % ng) * 4;
    float acc[4][4] = {};
#pragma unroll 4
    for (int d = 0; d < dh; d++) {
      const float4 g4 = *reinterpret_cast<const float4*>(sdOt + d * Np + i0);
      const float4 v4 = *reinterpret_cast<const float4*>(sVt + d * Np + j0);
      NT_FMA4x4(acc, g4, v4)
    }
#pragma unroll
    for (int r = 0; r < 4; r++)
      *reinterpret_cast<float4*>(sdS + (i0 + r) * Np + j0) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  // dS = P o (dP - rowsum(dP o P))   (attention.py:75, closed form of the per-row Jacobian); padded entries stay 0
  const int warp = tid >> 5, lane = tid & 31, nw = ATT_THREADS >> 5;
  for (int i = warp; i < N4; i += nw) {
    float s = 0.f;
    for (int j = lane; j < N4; j += 32) s += sdS[i * Np + j] * sP[i * Np + j];
    s = warp_sum(s);
    for (int j = lane; j < N4; j += 32) sdS[i * Np + j] = sP[i * Np + j] * (sdS[i * Np + j] - s);
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  float* obase = dqkv + (long long)b * N * 3 * D + h * dh;
  // phase A: dQ = dS K / sqrt(dh)  (attention.py:78), tile = 4 query rows x 4 features
  // phase B: dK = dS^T Q / sqrt(dh) (attention.py:79) and dV = P^T dO (attention.py:76), tile = 4 keys x 4 features
  const int itemsA = ng * dh4;
  for (int e = tid; e < 2 * itemsA; e += ATT_THREADS) {
    if (e < itemsA) {
      const int i0 = (e / dh4) * 4, d4 = (e % dh4) * 4;
      float acc[4][4] = {};
      const float* s0 = sdS + i0 * Np;
#pragma unroll 2
      for (int j = 0; j < N; j++) {
        const float4 s4 = make_float4(s0[j], s0[Np + j], s0[2 * Np + j], s0[3 * Np + j]);
        const float4 k4 = *reinterpret_cast<const float4*>(sK + j * dh + d4);
        NT_FMA4x4(acc, s4, k4)
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (i0 + r < N)
          *reinterpret_cast<float4*>(obase + (long long)(i0 + r) * 3 * D + d4) =
              make_float4(acc[r][0] * scale, acc[r][1] * scale, acc[r][2] * scale, acc[r][3] * scale);
    } else {
      const int f = e - itemsA;
      const int j0 = (f / dh4) * 4, d4 = (f % dh4) * 4;
      float ak[4][4] = {}, av[4][4] = {};
#pragma unroll 2
      for (int i = 0; i < N; i++) {
        const float4 s4 = *reinterpret_cast<const float4*>(sdS + i * Np + j0);
        const float4 p4 = *reinterpret_cast<const float4*>(sP + i * Np + j0);
        const float4 q4 = *reinterpret_cast<const float4*>(sQ + i * dh + d4);
        const float4 g4 = *reinterpret_cast<const float4*>(sdO + i * dh + d4);
        NT_FMA4x4(ak, s4, q4)
        NT_FMA4x4(av, p4, g4)
      }
#pragma unroll
      for (int r = 0; r < 4; r++)
        if (j0 + r < N) {
          float* o = obase + (long long)(j0 + r) * 3 * D + d4;
          *reinterpret_cast<float4*>(o + D) = make_float4(ak[r][0] * scale, ak[r][1] * scale, ak[r][2] * scale, ak[r][3] * scale);
          *reinterpret_cast<float4*>(o + 2 * D) = make_float4(av[r][0], av[r][1], av[r][2], av[r][3]);
        }
    }
  }
}

static size_t fwd_small_smem(int N, int dh) {
  const int N4 = (N + 3) & ~3, Np = N4 + 4;
  return sizeof(float) * (2 * (size_t)dh * Np + (size_t)N * dh + (size_t)N4 * Np);
}
static size_t bwd_small_smem(int N, int dh) {
  const int N4 = (N + 3) & ~3, Np = N4 + 4;
  return sizeof(float) * (3 * (size_t)N * dh + 2 * (size_t)dh * Np + 2 * (size_t)N4 * Np);
}
static bool small_ok(int N, int H, int dh) { return N <= 64 && dh <= 64 && (dh & 3) == 0; }

static int softmax_rows(const float* x, float* p, long long rows, int cols, const float* mask, int mask_kind, int N) {
  if (rows == 0) return 0;
  NT_CHECK(nt::launch_k(softmax_rows_kernel, dim3(ceil_div(rows * 32, 256)), dim3(256), 0, x, p, rows, cols, mask, mask_kind, N));
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
  NT_CHECK(nt::launch_k(softmax_bwd_rows_kernel, dim3(ceil_div((long long)rows * 32, 256)), dim3(256), 0, dp, p, dx, rows, cols));
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
  if (small_ok(N, H, dh)) {
    size_t smem = fwd_small_smem(N, dh);
#define NT_ATT_FWD(NC_, DHC_)                                                                                       \
    {                                                                                                               \
      static size_t configured = 0;                                                                                 \
      if (smem > 48 * 1024 && smem > configured) {                                                                  \
        NT_CHECK(cudaFuncSetAttribute(attn_fwd_small_kernel<NC_, DHC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                  \
        configured = smem;                                                                                          \
      }                                                                                                             \
      NT_CHECK(nt::launch_k(attn_fwd_small_kernel<NC_, DHC_>, dim3(B * H), dim3(ATT_THREADS), smem, qkv, mask, mask_kind, out, P, S, N, H, \
                                                                               dh, residual));                       \
    }
    if (N == 49 && dh == 24) NT_ATT_FWD(49, 24)
    else if (N == 49 && dh == 64) NT_ATT_FWD(49, 64)
    else NT_ATT_FWD(0, 0)
#undef NT_ATT_FWD
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
  if (small_ok(N, H, dh)) {
    size_t smem = bwd_small_smem(N, dh);
#define NT_ATT_BWD(NC_, DHC_)                                                                                       \
    {                                                                                                               \
      static size_t configured = 0;                                                                                 \
      if (smem > 48 * 1024 && smem > configured) {                                                                  \
        NT_CHECK(cudaFuncSetAttribute(attn_bwd_small_kernel<NC_, DHC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                      (int)smem));                                                                  \
        configured = smem;                                                                                          \
      }                                                                                                             \
      NT_CHECK(nt::launch_k(attn_bwd_small_kernel<NC_, DHC_>, dim3(B * H), dim3(ATT_THREADS), smem, dout, qkv, P, dqkv, N, H, dh));        \
    }
    if (N == 49 && dh == 24) NT_ATT_BWD(49, 24)
    else if (N == 49 && dh == 64) NT_ATT_BWD(49, 64)
    else NT_ATT_BWD(0, 0)
#undef NT_ATT_BWD
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
  NT_CHECK(nt::launch_k(softmax_bwd_rows_kernel, dim3(ceil_div((long long)B * H * N * 32, 256)), dim3(256), 0, scratch, P, scratch,
                                                                                         (long long)B * H * N, N));
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
