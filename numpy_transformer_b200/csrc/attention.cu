// Softmax rows and the attention core (attention.py:31-46,63-84; gpt/attdecoderblock.py:52-69,79-99).
// Two paths: one fused CTA per (sample, head) with everything in shared memory when N<=64 and
// dh<=64 (the ViT shapes), otherwise batched GEMMs + masked row softmax (the T=256 GPT shapes).
#include "common.cuh"
#include <cmath>
#include <cstdlib>

namespace nt {
bool attention_bf16_ok(int N, int H, int dh);
// attention_tc.cu: tcgen05 / TMEM kernels for dh == 64, 16 <= N <= 64 (no dense mask, no score output)
bool attention_tc_ok(int N, int H, int dh, int mask_kind, bool want_scores);
bool attention_tcl_ok(int N, int H, int dh, int mask_kind, bool want_scores);
int attention_fwd_tcl(const float* qkv, const void* qkv_hi, const void* qkv_lo, int mask_kind, float* out, float* P, int B, int N,
                      int H, const float* residual, void* o_hi, void* o_lo, float* lse);
bool attention_bwd_tcl_ok(int N, int H, int dh, int mask_kind);
int attention_bwd_tcl(const float* dout, const void* dout_hi, const void* dout_lo, const void* qkv_hi, const void* qkv_lo,
                      const float* att_out, const float* lse, float* dqkv, void* g_hi, void* g_lo, int B, int N, int H, int causal);
int attention_fwd_tc(const float* qkv, int mask_kind, float* out, float* P, int B, int N, int H, const float* residual, void* o_hi,
                     void* o_lo);
int attention_bwd_tc(const float* dout, const float* qkv, float* dqkv, int B, int N, int H, int causal, void* g_hi, void* g_lo);
int attention_fwd_bf16(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N, int H,
                       const float* residual, void* o_hi, void* o_lo);
int attention_bwd_bf16(const float* dout, const float* qkv, const float* P, float* dqkv, float* tsum, int B, int N, int H,
                       int causal, void* g_hi, void* g_lo, const float* att_out);

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
// ---------------------------------------------------------------- mma.sync (TF32 x3) small-N kernels
// N <= 64, dh in {24, 64}.  One CTA (4 warps) per (sample, head); warp w owns query rows 16w..16w+15 for
// S = QK^T, the softmax (on the accumulator fragments, in registers) and O = PV, so only the initial load
// needs a CTA barrier in forward.  Every operand is split hi/lo in registers (hi = top 19 bits, lo = x - hi)
// and each product is three m16n8k8 MMAs (lo*hi + hi*lo + hi*hi): fp32-class accuracy at ~1/10 of the
// FFMA instruction count.  Shared-memory row strides are chosen per access pattern: stride = 4 (mod 32) is
// conflict-free for (row=groupID, col=tig) fragment reads, stride = 8 (mod 32) for (row=tig, col=groupID).
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xFFFFE000u;
  lo = __float_as_uint(x - __uint_as_float(hi)) & 0xFFFFE000u;
}
template <int NA>
__device__ __forceinline__ void split_frag(const float (&x)[NA], uint32_t (&hi)[NA], uint32_t (&lo)[NA]) {
#pragma unroll
  for (int i = 0; i < NA; i++) split_tf32(x[i], hi[i], lo[i]);
}
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                     const uint32_t (&bh)[2], const uint32_t (&bl)[2]) {
  mma_tf32(d, al, bh);
  mma_tf32(d, ah, bl);
  mma_tf32(d, ah, bh);
}
// A fragment (16x8) of a row-major matrix M[row][col] with leading dimension ld, tile origin (r0, c0)
__device__ __forceinline__ void load_a(const float* M, int ld, int r0, int c0, int g, int t, float (&a)[4]) {
  a[0] = M[(r0 + g) * ld + c0 + t];
  a[1] = M[(r0 + g + 8) * ld + c0 + t];
  a[2] = M[(r0 + g) * ld + c0 + t + 4];
  a[3] = M[(r0 + g + 8) * ld + c0 + t + 4];
}
// A fragment of the TRANSPOSE of M: A(row, col) = M[col][row]
__device__ __forceinline__ void load_at(const float* M, int ld, int r0, int c0, int g, int t, float (&a)[4]) {
  a[0] = M[(c0 + t) * ld + r0 + g];
  a[1] = M[(c0 + t) * ld + r0 + g + 8];
  a[2] = M[(c0 + t + 4) * ld + r0 + g];
  a[3] = M[(c0 + t + 4) * ld + r0 + g + 8];
}
// B fragment (8x8, "col" layout) where B(k, n) = M[n][k]   (e.g. K in Q K^T)
__device__ __forceinline__ void load_b_nk(const float* M, int ld, int n0, int k0, int g, int t, float (&b)[2]) {
  b[0] = M[(n0 + g) * ld + k0 + t];
  b[1] = M[(n0 + g) * ld + k0 + t + 4];
}
// B fragment where B(k, n) = M[k][n]   (e.g. V in P V)
__device__ __forceinline__ void load_b_kn(const float* M, int ld, int k0, int n0, int g, int t, float (&b)[2]) {
  b[0] = M[(k0 + t) * ld + n0 + g];
  b[1] = M[(k0 + t + 4) * ld + n0 + g];
}

// asynchronous global->shared copies (LDGSTS): the whole operand set of a CTA is in flight at once instead of
// being serialised through registers (ncu: long-scoreboard stalls dominated the first version of these kernels)
__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int DH>
struct AttMma {
  static constexpr int LDA = DH + 4;                       // = 4 (mod 32) for DH in {24, 64}... (28 / 68)
  static constexpr int LDB = (DH % 32 == 0) ? DH + 8 : DH + 16;   // = 8 (mod 32): 72 / 40
  static constexpr int LDP = 68;                           // [64][68] score tiles
  static constexpr int KD = DH / 8;                        // k-steps over the head dimension
};

template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_mma_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                           int mask_kind, float* __restrict__ out, float* __restrict__ P,
                                                           float* __restrict__ S, int N, int H,
                                                           const float* __restrict__ residual) {
  pdl_prologue();
  using C = AttMma<DH>;
  extern __shared__ __align__(16) float sm[];
  float* sQ = sm;                          // [64][LDA]
  float* sK = sQ + 64 * C::LDA;            // [64][LDA]   B(k=d, n=j) = K[j][d]
  float* sV = sK + 64 * C::LDA;            // [64][LDB]   B(k=j, n=d) = V[j][d]
  float* sP = sV + 64 * C::LDB;            // [64][LDP]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  for (int e = tid; e < 64 * (DH / 4); e += 128) {
    const int i = e / (DH / 4), d4 = (e % (DH / 4)) * 4;
    if (i < N) {
      const float* r = base + (long long)i * 3 * D + d4;
      cp_async16(sQ + i * C::LDA + d4, r);
      cp_async16(sK + i * C::LDA + d4, r + D);
      cp_async16(sV + i * C::LDB + d4, r + 2 * D);
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sQ + i * C::LDA + d4) = z;
      *reinterpret_cast<float4*>(sK + i * C::LDA + d4) = z;
      *reinterpret_cast<float4*>(sV + i * C::LDB + d4) = z;
    }
  }
  cp_async_wait_all();
  __syncthreads();
  const int i0 = warp * 16;
  if (i0 >= N) return;                     // no further CTA-wide barriers below
  const int NT = (N + 7) >> 3;             // key tiles of 8
  float acc[8][4];
#pragma unroll
  for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < C::KD; ks++) {
    float a[4];
    uint32_t ah[4], al[4];
    load_a(sQ, C::LDA, i0, ks * 8, g, t, a);
    split_frag<4>(a, ah, al);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      if (nt < NT) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_nk(sK, C::LDA, nt * 8, ks * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(acc[nt], ah, al, bh_, bl_);
      }
    }
  }
  // scale, optional score dump, mask, softmax over the accumulator fragments (rows r0 = i0+g, r1 = r0+8)
  const float scale = 1.0f / sqrtf((float)DH);
  const int r0 = i0 + g, r1 = r0 + 8;
  float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int row = (c < 2) ? r0 : r1, col = nt * 8 + 2 * t + (c & 1);
      float v = acc[nt][c] * scale;                                           // attention.py:37
      if (nt < NT && col < N && row < N) {
        if (S) S[(long long)bh * N * N + row * N + col] = v;                   // att_col
        if (mask_kind == NT_MASK_CAUSAL) { if (col > row) v = -INFINITY; }
        else if (mask_kind == NT_MASK_DENSE) v += mask[row * N + col];         // attention.py:38-39
      } else {
        v = -INFINITY;                                                         // padded keys never win
      }
      acc[nt][c] = v;
      if (c < 2) mx0 = fmaxf(mx0, v); else mx1 = fmaxf(mx1, v);
    }
  }
  mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
  mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
  if (r0 >= N) mx0 = 0.f;                  // fully padded rows: keep the arithmetic finite
  if (r1 >= N) mx1 = 0.f;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    acc[nt][0] = expf(acc[nt][0] - mx0); acc[nt][1] = expf(acc[nt][1] - mx0);
    acc[nt][2] = expf(acc[nt][2] - mx1); acc[nt][3] = expf(acc[nt][3] - mx1);
    s0 += acc[nt][0] + acc[nt][1];
    s1 += acc[nt][2] + acc[nt][3];
  }
  s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
  s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
  const float inv0 = r0 < N ? 1.0f / s0 : 0.f, inv1 = r1 < N ? 1.0f / s1 : 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    const int col = nt * 8 + 2 * t;
    const float p00 = acc[nt][0] * inv0, p01 = acc[nt][1] * inv0, p10 = acc[nt][2] * inv1, p11 = acc[nt][3] * inv1;
    *reinterpret_cast<float2*>(sP + r0 * C::LDP + col) = make_float2(p00, p01);
    *reinterpret_cast<float2*>(sP + r1 * C::LDP + col) = make_float2(p10, p11);
    if (nt < NT) {
      float* pg0 = P + ((long long)bh * N + r0) * N + col;                     // atg__
      float* pg1 = P + ((long long)bh * N + r1) * N + col;
      if (r0 < N) { if (col < N) pg0[0] = p00; if (col + 1 < N) pg0[1] = p01; }
      if (r1 < N) { if (col < N) pg1[0] = p10; if (col + 1 < N) pg1[1] = p11; }
    }
  }
  __syncwarp();
  // O = P V : this warp's 16 rows x DH
  float o[C::KD][4];
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  for (int kt = 0; kt < NT; kt++) {
    float a[4];
    uint32_t ah[4], al[4];
    load_a(sP, C::LDP, i0, kt * 8, g, t, a);
    split_frag<4>(a, ah, al);
#pragma unroll
    for (int nd = 0; nd < C::KD; nd++) {
      float bb[2];
      uint32_t bh_[2], bl_[2];
      load_b_kn(sV, C::LDB, kt * 8, nd * 8, g, t, bb);
      split_frag<2>(bb, bh_, bl_);
      mma3(o[nd], ah, al, bh_, bl_);
    }
  }
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) {
    const int col = h * DH + nd * 8 + 2 * t;
    if (r0 < N) {
      const long long oi = ((long long)b * N + r0) * D + col;
      float2 v = make_float2(o[nd][0], o[nd][1]);
      if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
      *reinterpret_cast<float2*>(out + oi) = v;
    }
    if (r1 < N) {
      const long long oi = ((long long)b * N + r1) * D + col;
      float2 v = make_float2(o[nd][2], o[nd][3]);
      if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
      *reinterpret_cast<float2*>(out + oi) = v;
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_mma_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                           const float* __restrict__ P, float* __restrict__ dqkv, int N,
                                                           int H) {
  pdl_prologue();
  using C = AttMma<DH>;
  extern __shared__ __align__(16) float sm[];
  float* sQ = sm;                          // [64][LDB]  B(k=i, n=d) = Q[i][d]       (dK)
  float* sK = sQ + 64 * C::LDB;            // [64][LDB]  B(k=j, n=d) = K[j][d]       (dQ)
  float* sV = sK + 64 * C::LDB;            // [64][LDA]  B(k=d, n=j) = V[j][d]       (dP)
  float* sdO = sV + 64 * C::LDA;           // [64][LDA]  A in dP, B(k=i, n=d) in dV
  float* sP = sdO + 64 * C::LDA;           // [64][LDP]
  float* sdS = sP + 64 * C::LDP;           // [64][LDP]
  const int bh = blockIdx.x, b = bh / H, h = bh % H;
  const int D = H * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  for (int e = tid; e < 64 * (DH / 4); e += 128) {
    const int i = e / (DH / 4), d4 = (e % (DH / 4)) * 4;
    if (i < N) {
      const float* r = base + (long long)i * 3 * D + d4;
      cp_async16(sQ + i * C::LDB + d4, r);
      cp_async16(sK + i * C::LDB + d4, r + D);
      cp_async16(sV + i * C::LDA + d4, r + 2 * D);
      cp_async16(sdO + i * C::LDA + d4, dout + ((long long)b * N + i) * D + h * DH + d4);
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(sQ + i * C::LDB + d4) = z;
      *reinterpret_cast<float4*>(sK + i * C::LDB + d4) = z;
      *reinterpret_cast<float4*>(sV + i * C::LDA + d4) = z;
      *reinterpret_cast<float4*>(sdO + i * C::LDA + d4) = z;
    }
  }
  for (int e = tid; e < 64 * 64; e += 128) {
    const int i = e >> 6, j = e & 63;
    if (i < N && j < N) cp_async4(sP + i * C::LDP + j, P + (long long)bh * N * N + i * N + j);
    else sP[i * C::LDP + j] = 0.f;
  }
  cp_async_wait_all();
  __syncthreads();
  const int NT = (N + 7) >> 3;
  const int i0 = warp * 16;
  const float scale = 1.0f / sqrtf((float)DH);
  float* obase = dqkv + (long long)b * N * 3 * D + h * DH;
  const int r0 = i0 + g, r1 = r0 + 8;
  {
    // dP = dO V^T  (attention.py:74) for this warp's 16 query rows; rows >= N are all-zero and harmless
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KD; ks++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sdO, C::LDA, i0, ks * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        if (nt < NT) {
          float bb[2];
          uint32_t bh_[2], bl_[2];
          load_b_nk(sV, C::LDA, nt * 8, ks * 8, g, t, bb);
          split_frag<2>(bb, bh_, bl_);
          mma3(acc[nt], ah, al, bh_, bl_);
        }
      }
    }
    // dS = P o (dP - rowsum(dP o P))  (attention.py:75), on the fragments
    float s0 = 0.f, s1 = 0.f;
    float2 p0[8], p1[8];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      p0[nt] = *reinterpret_cast<const float2*>(sP + r0 * C::LDP + nt * 8 + 2 * t);
      p1[nt] = *reinterpret_cast<const float2*>(sP + r1 * C::LDP + nt * 8 + 2 * t);
      s0 += acc[nt][0] * p0[nt].x + acc[nt][1] * p0[nt].y;
      s1 += acc[nt][2] * p1[nt].x + acc[nt][3] * p1[nt].y;
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      *reinterpret_cast<float2*>(sdS + r0 * C::LDP + nt * 8 + 2 * t) =
          make_float2(p0[nt].x * (acc[nt][0] - s0), p0[nt].y * (acc[nt][1] - s0));
      *reinterpret_cast<float2*>(sdS + r1 * C::LDP + nt * 8 + 2 * t) =
          make_float2(p1[nt].x * (acc[nt][2] - s1), p1[nt].y * (acc[nt][3] - s1));
    }
  }
  __syncwarp();
  if (i0 < N) {
    // dQ = dS K / sqrt(dh)  (attention.py:78)
    float o[C::KD][4];
#pragma unroll
    for (int nd = 0; nd < C::KD; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
    for (int kt = 0; kt < NT; kt++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sdS, C::LDP, i0, kt * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nd = 0; nd < C::KD; nd++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_kn(sK, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(o[nd], ah, al, bh_, bl_);
      }
    }
#pragma unroll
    for (int nd = 0; nd < C::KD; nd++) {
      const int col = nd * 8 + 2 * t;
      if (r0 < N) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * D + col) = make_float2(o[nd][0] * scale, o[nd][1] * scale);
      if (r1 < N) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * D + col) = make_float2(o[nd][2] * scale, o[nd][3] * scale);
    }
  }
  __syncthreads();                         // every warp's rows of dS are in shared memory
  if (i0 < N) {
    // this warp now owns KEYS j0 = 16w..: dK = dS^T Q / sqrt(dh) (attention.py:79), dV = P^T dO (attention.py:76)
    const int j0 = i0;
    float ok[C::KD][4], ov[C::KD][4];
#pragma unroll
    for (int nd = 0; nd < C::KD; nd++) {
      ok[nd][0] = ok[nd][1] = ok[nd][2] = ok[nd][3] = 0.f;
      ov[nd][0] = ov[nd][1] = ov[nd][2] = ov[nd][3] = 0.f;
    }
    for (int kt = 0; kt < NT; kt++) {       // contraction over query rows i
      float a[4];
      uint32_t sh[4], sl[4], ph[4], pl[4];
      load_at(sdS, C::LDP, j0, kt * 8, g, t, a);
      split_frag<4>(a, sh, sl);
      load_at(sP, C::LDP, j0, kt * 8, g, t, a);
      split_frag<4>(a, ph, pl);
#pragma unroll
      for (int nd = 0; nd < C::KD; nd++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_kn(sQ, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(ok[nd], sh, sl, bh_, bl_);
        load_b_kn(sdO, C::LDA, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(ov[nd], ph, pl, bh_, bl_);
      }
    }
    const int k0r = j0 + g, k1r = k0r + 8;
#pragma unroll
    for (int nd = 0; nd < C::KD; nd++) {
      const int col = nd * 8 + 2 * t;
      if (k0r < N) {
        float* o = obase + (long long)k0r * 3 * D + col;
        *reinterpret_cast<float2*>(o + D) = make_float2(ok[nd][0] * scale, ok[nd][1] * scale);
        *reinterpret_cast<float2*>(o + 2 * D) = make_float2(ov[nd][0], ov[nd][1]);
      }
      if (k1r < N) {
        float* o = obase + (long long)k1r * 3 * D + col;
        *reinterpret_cast<float2*>(o + D) = make_float2(ok[nd][2] * scale, ok[nd][3] * scale);
        *reinterpret_cast<float2*>(o + 2 * D) = make_float2(ov[nd][2], ov[nd][3]);
      }
    }
  }
}

static bool g_att_mma_long = true;          // NTB200_ATT_LONG=0 falls back to batched SIMT GEMMs
// ---------------------------------------------------------------- mma.sync long-sequence kernels (GPT: T=256, dh=64)
// Same fragment machinery, tiled 64 queries x 64 keys.  The reference keeps the full softmax matrix P
// (atg__, gpt/attdecoderblock.py:64), so forward is two passes over the key chunks: pass 1 writes the scaled,
// masked scores into P and tracks the running row max / sum, pass 2 re-reads them, normalises in place and
// accumulates O = P V.  Backward is two kernels: per query block (dP -> row sums -> dS -> dQ, dS parked in the
// caller's scratch) and per key block (dK = dS^T Q, dV = P^T dO).  Causal masks skip the upper-triangular blocks.
constexpr int LQ = 64;
template <int DH>
__device__ __forceinline__ void load_rows_async(float* dst, int ld, const float* src, long long src_ld, int row0, int N,
                                                int tid) {
  // 64 rows x DH floats, rows >= N zero-filled
  for (int e = tid; e < LQ * (DH / 4); e += 128) {
    const int i = e / (DH / 4), d4 = (e % (DH / 4)) * 4;
    if (row0 + i < N) cp_async16(dst + i * ld + d4, src + (long long)(row0 + i) * src_ld + d4);
    else *reinterpret_cast<float4*>(dst + i * ld + d4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
// 64x64 tile of an [N][N] matrix (N % 4 == 0), origin (r0, c0), out-of-range entries zero-filled
__device__ __forceinline__ void load_tile_async(float* dst, int ld, const float* src, int N, int r0, int c0, int tid) {
  for (int e = tid; e < LQ * 16; e += 128) {
    const int i = e >> 4, c4 = (e & 15) * 4;
    if (r0 + i < N && c0 + c4 < N) cp_async16(dst + i * ld + c4, src + (long long)(r0 + i) * N + c0 + c4);
    else *reinterpret_cast<float4*>(dst + i * ld + c4) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_mma_long_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                int mask_kind, float* __restrict__ out, float* P,
                                                                float* __restrict__ S, int N, int H,
                                                                const float* __restrict__ residual) {
  pdl_prologue();
  using C = AttMma<DH>;
  extern __shared__ __align__(16) float sm[];
  float* sQ = sm;                          // [64][LDA]
  float* sK = sQ + 64 * C::LDA;            // [64][LDA]
  float* sV = sK + 64 * C::LDA;            // [64][LDB]
  float* sP = sV + 64 * C::LDB;            // [64][LDP]
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * LQ;
  const int D = H * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  float* Pb = P + (long long)bh * N * N;
  const int nchunks_all = (N + LQ - 1) / LQ;
  // the optional score dump (att_col) is PRE-mask, so it needs every block even under a causal mask
  const int nchunks = (mask_kind == NT_MASK_CAUSAL && S == nullptr) ? min(nchunks_all, (int)blockIdx.x + 1) : nchunks_all;
  load_rows_async<DH>(sQ, C::LDA, base, 3 * D, q0, N, tid);
  const int i0 = warp * 16;
  const int r0 = q0 + i0 + g, r1 = r0 + 8;
  const float scale = 1.0f / sqrtf((float)DH);
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  // ---------------- pass 1: scores -> P (raw), running max / sum
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_rows_async<DH>(sK, C::LDA, base + D, 3 * D, kc * LQ, N, tid);
    cp_async_wait_all();
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KD; ks++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sQ, C::LDA, i0, ks * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_nk(sK, C::LDA, nt * 8, ks * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(acc[nt], ah, al, bh_, bl_);
      }
    }
    float cm0 = -INFINITY, cm1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int row = (c < 2) ? r0 : r1, col = kc * LQ + nt * 8 + 2 * t + (c & 1);
        float v = acc[nt][c] * scale;                                          // attdecoderblock.py:57
        if (col < N && row < N) {
          if (S) S[(long long)bh * N * N + (long long)row * N + col] = v;       // att_col
          if (mask_kind == NT_MASK_CAUSAL) { if (col > row) v = -INFINITY; }
          else if (mask_kind == NT_MASK_DENSE) v += mask[(long long)row * N + col];
          Pb[(long long)row * N + col] = v;                                     // parked for pass 2
        } else {
          v = -INFINITY;
        }
        acc[nt][c] = v;
        if (c < 2) cm0 = fmaxf(cm0, v); else cm1 = fmaxf(cm1, v);
      }
    }
    cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 1)); cm0 = fmaxf(cm0, __shfl_xor_sync(0xffffffffu, cm0, 2));
    cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 1)); cm1 = fmaxf(cm1, __shfl_xor_sync(0xffffffffu, cm1, 2));
    const float n0_ = fmaxf(m0, cm0), n1_ = fmaxf(m1, cm1);
    float s0 = 0.f, s1 = 0.f;
    if (n0_ > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) s0 += expf(acc[nt][0] - n0_) + expf(acc[nt][1] - n0_);
    }
    if (n1_ > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) s1 += expf(acc[nt][2] - n1_) + expf(acc[nt][3] - n1_);
    }
    s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
    s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
    l0 = (n0_ > -INFINITY ? l0 * expf(m0 - n0_) : 0.f) + s0;
    l1 = (n1_ > -INFINITY ? l1 * expf(m1 - n1_) : 0.f) + s1;
    m0 = n0_;
    m1 = n1_;
  }
  const float inv0 = (r0 < N && l0 > 0.f) ? 1.0f / l0 : 0.f, inv1 = (r1 < N && l1 > 0.f) ? 1.0f / l1 : 0.f;
  if (!(m0 > -INFINITY)) m0 = 0.f;
  if (!(m1 > -INFINITY)) m1 = 0.f;
  // ---------------- pass 2: normalise P in place, O = P V
  float o[C::KD][4];
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_rows_async<DH>(sV, C::LDB, base + 2 * D, 3 * D, kc * LQ, N, tid);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
      float p00 = 0.f, p01 = 0.f, p10 = 0.f, p11 = 0.f;
      if (r0 < N) {
        if (col < N) { p00 = expf(Pb[(long long)r0 * N + col] - m0) * inv0; Pb[(long long)r0 * N + col] = p00; }
        if (col + 1 < N) { p01 = expf(Pb[(long long)r0 * N + col + 1] - m0) * inv0; Pb[(long long)r0 * N + col + 1] = p01; }
      }
      if (r1 < N) {
        if (col < N) { p10 = expf(Pb[(long long)r1 * N + col] - m1) * inv1; Pb[(long long)r1 * N + col] = p10; }
        if (col + 1 < N) { p11 = expf(Pb[(long long)r1 * N + col + 1] - m1) * inv1; Pb[(long long)r1 * N + col + 1] = p11; }
      }
      *reinterpret_cast<float2*>(sP + (i0 + g) * C::LDP + nt * 8 + 2 * t) = make_float2(p00, p01);
      *reinterpret_cast<float2*>(sP + (i0 + g + 8) * C::LDP + nt * 8 + 2 * t) = make_float2(p10, p11);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int kt = 0; kt < 8; kt++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sP, C::LDP, i0, kt * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nd = 0; nd < C::KD; nd++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_kn(sV, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(o[nd], ah, al, bh_, bl_);
      }
    }
  }
  // causal: the skipped upper-triangular blocks of P are exactly zero
  if (nchunks < nchunks_all) {
    for (int e = tid; e < LQ * (nchunks_all - nchunks) * LQ; e += 128) {
      const int i = e / ((nchunks_all - nchunks) * LQ), c = nchunks * LQ + e % ((nchunks_all - nchunks) * LQ);
      if (q0 + i < N && c < N) Pb[(long long)(q0 + i) * N + c] = 0.f;
    }
  }
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) {
    const int col = h * DH + nd * 8 + 2 * t;
    if (r0 < N) {
      const long long oi = ((long long)b * N + r0) * D + col;
      float2 v = make_float2(o[nd][0], o[nd][1]);
      if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
      *reinterpret_cast<float2*>(out + oi) = v;
    }
    if (r1 < N) {
      const long long oi = ((long long)b * N + r1) * D + col;
      float2 v = make_float2(o[nd][2], o[nd][3]);
      if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
      *reinterpret_cast<float2*>(out + oi) = v;
    }
  }
}

// backward, per query block: dS (parked in `scratch`) and dQ
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_mma_long_q_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                                  const float* __restrict__ P, float* __restrict__ dqkv,
                                                                  float* scratch, int N, int H, int causal) {
  pdl_prologue();
  using C = AttMma<DH>;
  extern __shared__ __align__(16) float sm[];
  float* sdO = sm;                         // [64][LDA]  A in dP
  float* sV = sdO + 64 * C::LDA;           // [64][LDA]  B(k=d, n=j) = V[j][d]
  float* sK = sV + 64 * C::LDA;            // [64][LDB]  B(k=j, n=d) = K[j][d]
  float* sdS = sK + 64 * C::LDB;           // [64][LDP]
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * LQ;
  const int D = H * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  const float* Pb = P + (long long)bh * N * N;
  float* Sb = scratch + (long long)bh * N * N;
  const int nchunks_all = (N + LQ - 1) / LQ;
  const int nchunks = causal ? min(nchunks_all, (int)blockIdx.x + 1) : nchunks_all;
  load_rows_async<DH>(sdO, C::LDA, dout + (long long)b * N * D + h * DH, D, q0, N, tid);
  const int i0 = warp * 16;
  const int r0 = q0 + i0 + g, r1 = r0 + 8;
  float d0 = 0.f, d1 = 0.f;                // rowsum(dP o P)
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_rows_async<DH>(sV, C::LDA, base + 2 * D, 3 * D, kc * LQ, N, tid);
    cp_async_wait_all();
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KD; ks++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sdO, C::LDA, i0, ks * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_nk(sV, C::LDA, nt * 8, ks * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(acc[nt], ah, al, bh_, bl_);
      }
    }
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int row = (c < 2) ? r0 : r1, cc = col + (c & 1);
        if (row < N && cc < N) {
          const long long idx = (long long)row * N + cc;
          Sb[idx] = acc[nt][c];                                    // dP parked for pass 2
          const float pv = Pb[idx];
          if (c < 2) d0 = fmaf(acc[nt][c], pv, d0); else d1 = fmaf(acc[nt][c], pv, d1);
        }
      }
    }
  }
  d0 += __shfl_xor_sync(0xffffffffu, d0, 1); d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
  d1 += __shfl_xor_sync(0xffffffffu, d1, 1); d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
  float o[C::KD][4];
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_rows_async<DH>(sK, C::LDB, base + D, 3 * D, kc * LQ, N, tid);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
      float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int row = (c < 2) ? r0 : r1, cc = col + (c & 1);
        if (row < N && cc < N) {
          const long long idx = (long long)row * N + cc;
          v[c] = Pb[idx] * (Sb[idx] - ((c < 2) ? d0 : d1));         // dS = P o (dP - rowsum)
          Sb[idx] = v[c];
        }
      }
      *reinterpret_cast<float2*>(sdS + (i0 + g) * C::LDP + nt * 8 + 2 * t) = make_float2(v[0], v[1]);
      *reinterpret_cast<float2*>(sdS + (i0 + g + 8) * C::LDP + nt * 8 + 2 * t) = make_float2(v[2], v[3]);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int kt = 0; kt < 8; kt++) {
      float a[4];
      uint32_t ah[4], al[4];
      load_a(sdS, C::LDP, i0, kt * 8, g, t, a);
      split_frag<4>(a, ah, al);
#pragma unroll
      for (int nd = 0; nd < C::KD; nd++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_kn(sK, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(o[nd], ah, al, bh_, bl_);
      }
    }
  }
  const float scale = 1.0f / sqrtf((float)DH);
  float* obase = dqkv + (long long)b * N * 3 * D + h * DH;
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) {
    const int col = nd * 8 + 2 * t;
    if (r0 < N) *reinterpret_cast<float2*>(obase + (long long)r0 * 3 * D + col) = make_float2(o[nd][0] * scale, o[nd][1] * scale);
    if (r1 < N) *reinterpret_cast<float2*>(obase + (long long)r1 * 3 * D + col) = make_float2(o[nd][2] * scale, o[nd][3] * scale);
  }
}

// backward, per key block: dK = dS^T Q / sqrt(dh), dV = P^T dO
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_mma_long_k_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                                  const float* __restrict__ P, float* __restrict__ dqkv,
                                                                  const float* __restrict__ scratch, int N, int H, int causal) {
  pdl_prologue();
  using C = AttMma<DH>;
  extern __shared__ __align__(16) float sm[];
  float* sQ = sm;                          // [64][LDB]  B(k=i, n=d)
  float* sdO = sQ + 64 * C::LDB;           // [64][LDB]  B(k=i, n=d)
  float* sP = sdO + 64 * C::LDB;           // [64 q][LDP]  read transposed
  float* sdS = sP + 64 * C::LDP;           // [64 q][LDP]
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * LQ;
  const int D = H * DH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  const float* Pb = P + (long long)bh * N * N;
  const float* Sb = scratch + (long long)bh * N * N;
  const int nchunks_all = (N + LQ - 1) / LQ;
  const int qc_begin = causal ? blockIdx.x : 0;
  const int j0 = warp * 16;
  float ok[C::KD][4], ov[C::KD][4];
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) {
    ok[nd][0] = ok[nd][1] = ok[nd][2] = ok[nd][3] = 0.f;
    ov[nd][0] = ov[nd][1] = ov[nd][2] = ov[nd][3] = 0.f;
  }
  for (int qc = qc_begin; qc < nchunks_all; qc++) {
    __syncthreads();
    load_rows_async<DH>(sQ, C::LDB, base, 3 * D, qc * LQ, N, tid);
    load_rows_async<DH>(sdO, C::LDB, dout + (long long)b * N * D + h * DH, D, qc * LQ, N, tid);
    load_tile_async(sP, C::LDP, Pb, N, qc * LQ, k0, tid);
    load_tile_async(sdS, C::LDP, Sb, N, qc * LQ, k0, tid);
    cp_async_wait_all();
    __syncthreads();
    for (int kt = 0; kt < 8; kt++) {       // contraction over the 64 query rows of this chunk
      float a[4];
      uint32_t sh[4], sl[4], ph[4], pl[4];
      load_at(sdS, C::LDP, j0, kt * 8, g, t, a);
      split_frag<4>(a, sh, sl);
      load_at(sP, C::LDP, j0, kt * 8, g, t, a);
      split_frag<4>(a, ph, pl);
#pragma unroll
      for (int nd = 0; nd < C::KD; nd++) {
        float bb[2];
        uint32_t bh_[2], bl_[2];
        load_b_kn(sQ, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(ok[nd], sh, sl, bh_, bl_);
        load_b_kn(sdO, C::LDB, kt * 8, nd * 8, g, t, bb);
        split_frag<2>(bb, bh_, bl_);
        mma3(ov[nd], ph, pl, bh_, bl_);
      }
    }
  }
  const float scale = 1.0f / sqrtf((float)DH);
  float* obase = dqkv + (long long)b * N * 3 * D + h * DH;
  const int kr0 = k0 + j0 + g, kr1 = kr0 + 8;
#pragma unroll
  for (int nd = 0; nd < C::KD; nd++) {
    const int col = nd * 8 + 2 * t;
    if (kr0 < N) {
      float* o = obase + (long long)kr0 * 3 * D + col;
      *reinterpret_cast<float2*>(o + D) = make_float2(ok[nd][0] * scale, ok[nd][1] * scale);
      *reinterpret_cast<float2*>(o + 2 * D) = make_float2(ov[nd][0], ov[nd][1]);
    }
    if (kr1 < N) {
      float* o = obase + (long long)kr1 * 3 * D + col;
      *reinterpret_cast<float2*>(o + D) = make_float2(ok[nd][2] * scale, ok[nd][3] * scale);
      *reinterpret_cast<float2*>(o + 2 * D) = make_float2(ov[nd][2], ov[nd][3]);
    }
  }
}

static bool long_ok(int N, int dh) { return g_att_mma_long && N > 64 && dh == 64 && (N % 4) == 0; }

template <int DH>
static size_t fwd_mma_smem() { using C = AttMma<DH>; return sizeof(float) * (64 * C::LDA * 2 + 64 * C::LDB + 64 * C::LDP); }
template <int DH>
static size_t bwd_mma_smem() { using C = AttMma<DH>; return sizeof(float) * (64 * C::LDB * 2 + 64 * C::LDA * 2 + 2 * 64 * C::LDP); }
static bool g_att_mma = true;               // NTB200_ATT_MMA=0 falls back to the FFMA kernels
static bool mma_ok(int N, int dh) { return g_att_mma && N <= 64 && (dh == 24 || dh == 64); }

static bool small_ok(int N, int H, int dh) { return N <= 64 && dh <= 64 && (dh & 3) == 0; }

static int softmax_rows(const float* x, float* p, long long rows, int cols, const float* mask, int mask_kind, int N) {
  if (rows == 0) return 0;
  NT_CHECK(nt::launch_k(softmax_rows_kernel, dim3(ceil_div(rows * 32, 256)), dim3(256), 0, x, p, rows, cols, mask, mask_kind, N));
  NT_LAUNCH_OK();
  return 0;
}

// ---------------------------------------------------------------- KV-cache decode step
// The reference regenerates by re-running the whole window for every new token (gpt_predict_poetrythree.py:108-121,
// gpt_train_potry3000.py:300-342).  While the window is still growing the keys / values of the earlier tokens do not
// change (causal mask, learned positions fixed per index), so they are cached per block and one step only attends
// the new token's query over them.
// `d_pos` (may be NULL): position read from device memory, so one captured graph serves every decode step.
__global__ void kv_append_kernel(const float* __restrict__ qkv, float* __restrict__ kc, float* __restrict__ vc, int B, int T,
                                 int Tmax, int pos0, int D, const int32_t* __restrict__ d_pos) {
  pdl_prologue();
  if (d_pos) pos0 = *d_pos;
  if (pos0 + T > Tmax) return;                          // full window: the host falls back to the sliding recompute
  const long long total = (long long)B * T * D;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(i % D);
    const long long bt = i / D;
    const int t = (int)(bt % T), b = (int)(bt / T);
    const float* row = qkv + ((long long)b * T + t) * 3 * D;
    const long long o = ((long long)b * Tmax + pos0 + t) * D + d;
    kc[o] = row[D + d];
    vc[o] = row[2 * D + d];
  }
}

// one 4-warp CTA per (sample, head): scores over the cached keys (one key per thread, float4 loads), softmax, then the
// weighted sum of the cached values (warp w takes keys w, w+4, ...; lanes run along dh so every load is one 128-byte line)
constexpr int DEC_THREADS = 128;
// `append`: the CTA first stores the new token's k / v slice of its head into cache row len-1 (nt_kv_append folded in; the
// caches are therefore plain pointers: the rows are re-read by the same CTA after the barrier).
__global__ void __launch_bounds__(DEC_THREADS) attn_decode_kernel(const float* __restrict__ qkv_new, float* kc, float* vc,
                                                                  float* __restrict__ out, int Tmax, int len, int H, int dh,
                                                                  const int32_t* __restrict__ d_pos, int append) {
  pdl_prologue();
  if (d_pos) len = min(*d_pos + 1, Tmax);
  extern __shared__ __align__(16) float sm[];          // [dh] q | [4][dh] partial outputs | [8] reductions | [len] p
  float* sq = sm;
  float* so = sm + dh;
  float* sr = so + 4 * dh;
  float* sp = sr + 8;
  const int bh = blockIdx.x, b = bh / H, h = bh % H, D = H * dh, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int d = tid; d < dh; d += DEC_THREADS) {
    const long long qo = (long long)b * 3 * D + h * dh + d;
    sq[d] = qkv_new[qo];
    if (append) {
      const long long co = ((long long)b * Tmax + (len - 1)) * D + h * dh + d;
      kc[co] = qkv_new[qo + D];
      vc[co] = qkv_new[qo + 2 * D];
    }
  }
  __syncthreads();
  const float scale = 1.0f / sqrtf((float)dh);
  const float* kb = kc + (long long)b * Tmax * D + h * dh;
  const float* vb = vc + (long long)b * Tmax * D + h * dh;
  const bool vec = (dh % 4 == 0) && (D % 4 == 0);
  float mx = -INFINITY;
  for (int j = tid; j < len; j += DEC_THREADS) {
    const float* kr = kb + (long long)j * D;
    float s = 0.f;
    if (vec) {
      for (int d = 0; d < dh; d += 4) {
        const float4 k4 = *reinterpret_cast<const float4*>(kr + d);
        const float4 q4 = *reinterpret_cast<const float4*>(sq + d);
        s = fmaf(q4.x, k4.x, s); s = fmaf(q4.y, k4.y, s); s = fmaf(q4.z, k4.z, s); s = fmaf(q4.w, k4.w, s);
      }
    } else {
      for (int d = 0; d < dh; d++) s = fmaf(sq[d], kr[d], s);
    }
    s *= scale;                                       // attdecoderblock.py:57; the new token sees every cached key
    sp[j] = s;
    mx = fmaxf(mx, s);
  }
  mx = warp_max(mx);
  if (lane == 0) sr[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(sr[0], sr[1]), fmaxf(sr[2], sr[3]));
  float sum = 0.f;
  for (int j = tid; j < len; j += DEC_THREADS) { const float e = expf(sp[j] - mx); sp[j] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) sr[4 + warp] = sum;
  __syncthreads();
  const float inv = 1.0f / (sr[4] + sr[5] + sr[6] + sr[7]);
  for (int d = lane; d < dh; d += 32) {
    float acc = 0.f;
    for (int j = warp; j < len; j += 4) acc = fmaf(sp[j], vb[(long long)j * D + d], acc);
    so[warp * dh + d] = acc;
  }
  __syncthreads();
  for (int d = tid; d < dh; d += DEC_THREADS)
    out[(long long)b * D + h * dh + d] = (so[d] + so[dh + d] + so[2 * dh + d] + so[3 * dh + d]) * inv;
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
  NT_FUSE_TRY(nt::fuse_attn_fwd(qkv, mask_kind == NT_MASK_DENSE ? mask : nullptr, mask_kind, out, P, S, B, N, H, dh, residual));
  const int D = H * dh;
  if (nt::attention_tc_ok(N, H, dh, mask_kind, S != nullptr)) return nt::attention_fwd_tc(qkv, mask_kind, out, P, B, N, H, residual, nullptr, nullptr);
  if (nt::attention_tcl_ok(N, H, dh, mask_kind, S != nullptr)) return nt::attention_fwd_tcl(qkv, nullptr, nullptr, mask_kind, out, P, B, N, H, residual, nullptr, nullptr, nullptr);
  if (nt::attention_bf16_ok(N, H, dh)) return nt::attention_fwd_bf16(qkv, mask, mask_kind, out, P, S, B, N, H, residual, nullptr, nullptr);
  if (const char* e = getenv("NTB200_ATT_MMA")) g_att_mma = (e[0] != '0');
  if (mma_ok(N, dh)) {
    if (dh == 24) {
      const size_t smem = fwd_mma_smem<24>();
      static bool cfg = false;
      if (!cfg) { NT_CHECK(cudaFuncSetAttribute(attn_fwd_mma_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
      NT_CHECK(nt::launch_k(attn_fwd_mma_kernel<24>, dim3(B * H), dim3(128), smem, qkv, mask, mask_kind, out, P, S, N, H, residual));
    } else {
      const size_t smem = fwd_mma_smem<64>();
      static bool cfg = false;
      if (!cfg) { NT_CHECK(cudaFuncSetAttribute(attn_fwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
      NT_CHECK(nt::launch_k(attn_fwd_mma_kernel<64>, dim3(B * H), dim3(128), smem, qkv, mask, mask_kind, out, P, S, N, H, residual));
    }
    NT_LAUNCH_OK();
    return 0;
  }
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
  if (const char* e = getenv("NTB200_ATT_LONG")) g_att_mma_long = (e[0] != '0');
  if (long_ok(N, dh)) {
    using C = AttMma<64>;
    const size_t smem = sizeof(float) * (64 * C::LDA * 2 + 64 * C::LDB + 64 * C::LDP);
    static bool cfg = false;
    if (!cfg) { NT_CHECK(cudaFuncSetAttribute(attn_fwd_mma_long_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
    NT_CHECK(nt::launch_k(attn_fwd_mma_long_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), smem, qkv, mask, mask_kind, out, P, S, N, H, residual));
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
                     int H, int dh, int mask_kind) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0, "nt_attention_bwd: bad shape");
  if (B == 0) return 0;
  NT_FUSE_TRY(nt::fuse_attn_bwd(dout, qkv, P, dqkv, B, N, H, dh));
  const int D = H * dh;
  // BF16x3 kernels (attention_bf16.cu): `scratch` only carries the B*H*N softmax-backward row sums
  if (nt::attention_tc_ok(N, H, dh, mask_kind, false))
    return nt::attention_bwd_tc(dout, qkv, dqkv, B, N, H, mask_kind == NT_MASK_CAUSAL ? 1 : 0, nullptr, nullptr);
  if ((scratch || N <= 64) && nt::attention_bf16_ok(N, H, dh))
    return nt::attention_bwd_bf16(dout, qkv, P, dqkv, scratch, B, N, H, mask_kind == NT_MASK_CAUSAL ? 1 : 0, nullptr, nullptr, nullptr);
  if (mma_ok(N, dh)) {
    if (dh == 24) {
      const size_t smem = bwd_mma_smem<24>();
      static bool cfg = false;
      if (!cfg) { NT_CHECK(cudaFuncSetAttribute(attn_bwd_mma_kernel<24>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
      NT_CHECK(nt::launch_k(attn_bwd_mma_kernel<24>, dim3(B * H), dim3(128), smem, dout, qkv, P, dqkv, N, H));
    } else {
      const size_t smem = bwd_mma_smem<64>();
      static bool cfg = false;
      if (!cfg) { NT_CHECK(cudaFuncSetAttribute(attn_bwd_mma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); cfg = true; }
      NT_CHECK(nt::launch_k(attn_bwd_mma_kernel<64>, dim3(B * H), dim3(128), smem, dout, qkv, P, dqkv, N, H));
    }
    NT_LAUNCH_OK();
    return 0;
  }
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
  if (long_ok(N, dh)) {
    using C = AttMma<64>;
    const int causal = (mask_kind == NT_MASK_CAUSAL) ? 1 : 0;
    const size_t smem_q = sizeof(float) * (64 * C::LDA * 2 + 64 * C::LDB + 64 * C::LDP);
    const size_t smem_k = sizeof(float) * (64 * C::LDB * 2 + 64 * C::LDP * 2);
    static bool cfg = false;
    if (!cfg) {
      NT_CHECK(cudaFuncSetAttribute(attn_bwd_mma_long_q_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_q));
      NT_CHECK(cudaFuncSetAttribute(attn_bwd_mma_long_k_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_k));
      cfg = true;
    }
    NT_CHECK(nt::launch_k(attn_bwd_mma_long_q_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), smem_q, dout, qkv, P, dqkv, scratch, N, H, causal));
    NT_LAUNCH_OK();
    NT_CHECK(nt::launch_k(attn_bwd_mma_long_k_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), smem_k, dout, qkv, P, dqkv, (const float*)scratch, N, H, causal));
    NT_LAUNCH_OK();
    return 0;
  }
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

int nt_attention_fwd_split(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N,
                           int H, int dh, const float* residual, void* out_hi, void* out_lo) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0, "nt_attention_fwd_split: bad shape");
  NT_REQUIRE(mask_kind != NT_MASK_DENSE || mask, "nt_attention_fwd_split: dense mask_kind needs a mask");
  NT_REQUIRE(nt::attention_bf16_ok(N, H, dh), "nt_attention_fwd_split: only the BF16x3 kernels (dh == 64, GEMM mode 1) emit split outputs");
  NT_REQUIRE((out_hi != nullptr) == (out_lo != nullptr), "nt_attention_fwd_split: pass both split planes or neither");
  if (B == 0) return 0;
  if (nt::attention_tc_ok(N, H, dh, mask_kind, S != nullptr)) return nt::attention_fwd_tc(qkv, mask_kind, out, P, B, N, H, residual, out_hi, out_lo);
  if (nt::attention_tcl_ok(N, H, dh, mask_kind, S != nullptr)) return nt::attention_fwd_tcl(qkv, nullptr, nullptr, mask_kind, out, P, B, N, H, residual, out_hi, out_lo, nullptr);
  return nt::attention_fwd_bf16(qkv, mask, mask_kind, out, P, S, B, N, H, residual, out_hi, out_lo);
}

int nt_attention_fwd_planes(const float* qkv, const void* qkv_hi, const void* qkv_lo, int mask_kind, float* out, float* P, int B,
                            int N, int H, int dh, const float* residual, void* out_hi, void* out_lo, float* lse, int* taken) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0 && taken, "nt_attention_fwd_planes: bad shape");
  NT_REQUIRE(qkv_hi && qkv_lo, "nt_attention_fwd_planes: needs both planes of qkv");
  NT_REQUIRE((out_hi != nullptr) == (out_lo != nullptr), "nt_attention_fwd_planes: pass both split planes or neither");
  *taken = 0;
  if (!nt::attention_tcl_ok(N, H, dh, mask_kind, false)) return 0;
  *taken = 1;
  if (B == 0) return 0;
  return nt::attention_fwd_tcl(qkv, qkv_hi, qkv_lo, mask_kind, out, P, B, N, H, residual, out_hi, out_lo, lse);
}

int nt_attention_bwd_planes(const float* dout, const void* dout_hi, const void* dout_lo, const void* qkv_hi, const void* qkv_lo,
                            const float* att_out, const float* lse, float* dqkv, void* dqkv_hi, void* dqkv_lo, int B, int N, int H,
                            int dh, int mask_kind, int* taken) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0 && taken, "nt_attention_bwd_planes: bad shape");
  NT_REQUIRE(dout && dout_hi && dout_lo && qkv_hi && qkv_lo && att_out && lse && dqkv, "nt_attention_bwd_planes: missing operand");
  NT_REQUIRE((dqkv_hi != nullptr) == (dqkv_lo != nullptr), "nt_attention_bwd_planes: pass both split planes or neither");
  *taken = 0;
  if (!nt::attention_bwd_tcl_ok(N, H, dh, mask_kind)) return 0;
  *taken = 1;
  if (B == 0) return 0;
  return nt::attention_bwd_tcl(dout, dout_hi, dout_lo, qkv_hi, qkv_lo, att_out, lse, dqkv, dqkv_hi, dqkv_lo, B, N, H,
                               mask_kind == NT_MASK_CAUSAL ? 1 : 0);
}

int nt_attention_bwd_split(const float* dout, const float* qkv, const float* P, float* dqkv, float* scratch, int B, int N,
                           int H, int dh, int mask_kind, void* dqkv_hi, void* dqkv_lo, const float* att_out) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && N > 0 && H > 0 && dh > 0, "nt_attention_bwd_split: bad shape");
  NT_REQUIRE(nt::attention_bf16_ok(N, H, dh) && (scratch || N <= 64),
             "nt_attention_bwd_split: only the BF16x3 kernels (dh == 64, GEMM mode 1, scratch of B*H*N floats) emit split outputs");
  NT_REQUIRE((dqkv_hi != nullptr) == (dqkv_lo != nullptr), "nt_attention_bwd_split: pass both split planes or neither");
  if (B == 0) return 0;
  if (nt::attention_tc_ok(N, H, dh, mask_kind, false))
    return nt::attention_bwd_tc(dout, qkv, dqkv, B, N, H, mask_kind == NT_MASK_CAUSAL ? 1 : 0, dqkv_hi, dqkv_lo);
  return nt::attention_bwd_bf16(dout, qkv, P, dqkv, scratch, B, N, H, mask_kind == NT_MASK_CAUSAL ? 1 : 0, dqkv_hi, dqkv_lo, att_out);
}

int nt_kv_append(const float* qkv, float* kcache, float* vcache, int B, int T, int Tmax, int pos0, int D) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && T >= 0 && pos0 >= 0 && pos0 + T <= Tmax && D > 0, "nt_kv_append: rows %d..%d do not fit a cache of %d", pos0, pos0 + T, Tmax);
  const long long total = (long long)B * T * D;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(kv_append_kernel, dim3((unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184)), dim3(256), 0, qkv, kcache,
                        vcache, B, T, Tmax, pos0, D, (const int32_t*)nullptr));
  NT_LAUNCH_OK();
  return 0;
}

int nt_kv_append_at(const float* qkv, float* kcache, float* vcache, int B, int Tmax, int D, const int32_t* d_pos) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && Tmax > 0 && D > 0 && d_pos, "nt_kv_append_at: bad arguments");
  const long long total = (long long)B * D;
  if (total == 0) return 0;
  NT_CHECK(nt::launch_k(kv_append_kernel, dim3((unsigned)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184)), dim3(256), 0, qkv, kcache,
                        vcache, B, 1, Tmax, 0, D, d_pos));
  NT_LAUNCH_OK();
  return 0;
}

int nt_attention_decode(const float* qkv_new, const float* kcache, const float* vcache, float* out, int B, int Tmax, int len,
                        int H, int dh) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && len >= 1 && len <= Tmax && H > 0 && dh > 0, "nt_attention_decode: bad shape (len %d, cache %d)", len, Tmax);
  if (B == 0) return 0;
  const size_t smem = sizeof(float) * (5 * (size_t)dh + 8 + len);
  NT_REQUIRE(smem <= 48 * 1024, "nt_attention_decode: context %d too long", len);
  NT_CHECK(nt::launch_k(attn_decode_kernel, dim3(B * H), dim3(DEC_THREADS), smem, qkv_new, const_cast<float*>(kcache),
                        const_cast<float*>(vcache), out, Tmax, len, H, dh, (const int32_t*)nullptr, 0));
  NT_LAUNCH_OK();
  return 0;
}

int nt_attention_decode_at(const float* qkv_new, float* kcache, float* vcache, float* out, int B, int Tmax, int H,
                           int dh, const int32_t* d_pos, int append) {
  NT_ENTRY();
  NT_REQUIRE(B >= 0 && Tmax >= 1 && H > 0 && dh > 0 && d_pos, "nt_attention_decode_at: bad arguments");
  if (B == 0) return 0;
  NT_FUSE_TRY(nt::fuse_attn_decode(qkv_new, kcache, vcache, out, B, Tmax, H, dh, d_pos, append));
  const size_t smem = sizeof(float) * (5 * (size_t)dh + 8 + Tmax);
  NT_REQUIRE(smem <= 48 * 1024, "nt_attention_decode_at: context %d too long", Tmax);
  NT_CHECK(nt::launch_k(attn_decode_kernel, dim3(B * H), dim3(DEC_THREADS), smem, qkv_new, kcache, vcache, out, Tmax, 0, H, dh, d_pos,
                        append ? 1 : 0));
  NT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
