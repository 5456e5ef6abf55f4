// Attention core on the BF16x3 mma path (attention.py:31-46,63-84; gpt/attdecoderblock.py:52-69,79-99) for dh = 64,
// any sequence length (the scaled ViT: N = 49; the GPT configs: T = 256).
//
// Same arithmetic contract as the TF32x3 kernels of attention.cu (fp32-class: every product is lo*hi + hi*lo + hi*hi
// of bf16 splits, fp32 accumulate), but m16n8k16 MMAs fed by ldmatrix from operands that are split ONCE when a tile
// is staged in shared memory — half the tensor instructions and no per-fragment conversions.
//   forward   one CTA per (64-query block, sample, head), 4 warps x 16 query rows.  Pass 1 sweeps the key blocks:
//             scores, row max / sum, scores parked in the P buffer; pass 2 re-reads them (same thread, L2-resident),
//             writes the normalised softmax P (the reference keeps it: atg__) and accumulates O = P V with P fed
//             from the accumulator fragments.  (The first version recomputed Q K^T in pass 2; the kernel is bound by
//             shared-memory operand traffic, not by the tensor pipe, so the round trip is cheaper.)
//   backward  kernel Q per query block: t_i = sum_j dP_ij P_ij, then dS = P o (dP - t), dQ = dS K / sqrt(dh);
//             kernel K per key block: recomputes dP = dO V^T and dS for its keys, dK = dS^T Q / sqrt(dh),
//             dV = P^T dO.  Only the row sums t travel between the two (first B*H*N floats of `scratch`);
//             the B*H*N*N dS scratch of the TF32 path is gone.
// Causal masks skip the upper-triangular blocks (P is exactly 0 there).
#include "mma_bf16x3.cuh"
#include <cstdlib>

namespace nt {
namespace ab {
using namespace vb;

constexpr int LQ = 64;

// stage rows row0.. of a [N][ld] fp32 matrix (64 x DH tile) as a split operand; rows >= N are zero
template <int DH>
__device__ __forceinline__ void load_split(const SArr& dst, const float* src, long long ld, int row0, int N, int tid) {
  constexpr int IT = LQ * (DH / 4) / 128;
  float4 v[IT];
  // all loads of the tile first (the shared stores go through a generic pointer, so the compiler would otherwise
  // keep every load behind the previous iteration's stores: one full memory latency per 16 bytes)
#pragma unroll
  for (int i = 0; i < IT; i++) {
    const int e = tid + 128 * i, r = e / (DH / 4), c = (e - r * (DH / 4)) * 4;
    v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < N) v[i] = *reinterpret_cast<const float4*>(src + (long long)(row0 + r) * ld + c);
  }
#pragma unroll
  for (int i = 0; i < IT; i++) {
    const int e = tid + 128 * i, r = e / (DH / 4), c = (e - r * (DH / 4)) * 4;
    st2(dst, r, c, v[i].x, v[i].y);
    st2(dst, r, c + 2, v[i].z, v[i].w);
  }
}

// dqkv element pair -> fp32 and, when wanted, the bf16 hi / lo planes (same [B,N,3D] layout)
__device__ __forceinline__ void store_g(float* base, uint32_t* g_hi, uint32_t* g_lo, const float* dq0, long long off, float a, float b) {
  *reinterpret_cast<float2*>(base + off) = make_float2(a, b);
  if (g_hi) {
    uint32_t hh, ll;
    split2(a, b, hh, ll);
    const long long e = (base + off) - dq0;
    g_hi[e >> 1] = hh;
    g_lo[e >> 1] = ll;
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

template <int DH>
struct Smem {
  static constexpr int LD = DH + 8;
  static constexpr int ARR = LQ * LD * 2 * 2;          // one [64][DH] hi/lo pair
  static constexpr int PARR = LQ * LDP * 2 * 2;        // one [64][64] hi/lo pair
  static constexpr int PW = 16 * LDP * 2 * 2;          // one warp's [16][64] pair
};

// scaled + masked scores of this warp's 16 rows against the 64 keys of chunk kc, in fragment layout
template <int DH>
__device__ __forceinline__ void scores(float (&s)[8][4], const SArr& sQ, const SArr& sK, int i0, int r0, int r1, int kc,
                                       int N, float scale, int mask_kind, const float* __restrict__ mask, float* Sout,
                                       int lane, int t) {
#pragma unroll
  for (int nt = 0; nt < 8; nt++) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
  gemm_rowk_nk<DH>(s, sQ, i0, sK, 0, 4, lane);
#pragma unroll
  for (int nt = 0; nt < 8; nt++)
#pragma unroll
    for (int c = 0; c < 4; c++) {
      const int row = (c < 2) ? r0 : r1, col = kc * LQ + nt * 8 + 2 * t + (c & 1);
      float v = s[nt][c] * scale;                                             // attention.py:37 / attdecoderblock.py:57
      if (col < N && row < N) {
        if (Sout) Sout[(long long)row * N + col] = v;                          // att_col (pre-mask)
        if (mask_kind == NT_MASK_CAUSAL) { if (col > row) v = -INFINITY; }
        else if (mask_kind == NT_MASK_DENSE) v += mask[(long long)row * N + col];   // added AFTER the scale
      } else {
        v = -INFINITY;
      }
      s[nt][c] = v;
    }
}

template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_bf16_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                            int mask_kind, float* __restrict__ out, float* __restrict__ P,
                                                            float* __restrict__ S, int N, int H,
                                                            const float* __restrict__ residual, uint32_t* __restrict__ o_hi,
                                                            uint32_t* __restrict__ o_lo) {
  pdl_prologue();
  using M = Smem<DH>;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const SArr sQ = make_arr(smem, M::LD), sK = make_arr(smem + M::ARR, M::LD), sV = make_arr(smem + 2 * M::ARR, M::LD);
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * LQ, D = H * DH;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  float* Pb = P + (long long)bh * N * N;
  float* Sb = S ? S + (long long)bh * N * N : nullptr;
  const int nall = (N + LQ - 1) / LQ;
  const int nchunks = (mask_kind == NT_MASK_CAUSAL) ? min(nall, (int)blockIdx.x + 1) : nall;
  const int i0 = warp * 16, r0 = q0 + i0 + g, r1 = r0 + 8;
  const float scale = rsqrtf((float)DH);
  load_split<DH>(sQ, base, 3 * D, q0, N, tid);
  // ---------------- pass 1: scores once; row max / sum over the key blocks.  The scaled + masked scores are parked in
  // the P buffer (each thread re-reads exactly the elements it wrote, so no synchronisation is needed) instead of being
  // recomputed in pass 2: ncu showed the kernel bound by shared-memory operand traffic (L1 75 % busy, tensor pipe 24 %),
  // and the second Q K^T cost a K-tile staging + 20 KB of fragment loads per warp and block; the round trip through P
  // costs 2 x 16 KB of L2-resident global traffic instead.  The optional pre-mask score dump (att_col is dense) needs
  // every block even under a causal mask.
  float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
  const int npass1 = Sb ? nall : nchunks;
  for (int kc = 0; kc < npass1; kc++) {
    __syncthreads();
    load_split<DH>(sK, base + D, 3 * D, kc * LQ, N, tid);
    __syncthreads();
    float s[8][4];
    scores<DH>(s, sQ, sK, i0, r0, r1, kc, N, scale, mask_kind, mask, Sb, lane, t);
    if (kc >= nchunks) continue;              // beyond the causal frontier: only the score dump was needed
    float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
      if (r0 < N) { if (col < N) Pb[(long long)r0 * N + col] = s[nt][0]; if (col + 1 < N) Pb[(long long)r0 * N + col + 1] = s[nt][1]; }
      if (r1 < N) { if (col < N) Pb[(long long)r1 * N + col] = s[nt][2]; if (col + 1 < N) Pb[(long long)r1 * N + col + 1] = s[nt][3]; }
      c0 = fmaxf(c0, fmaxf(s[nt][0], s[nt][1]));
      c1 = fmaxf(c1, fmaxf(s[nt][2], s[nt][3]));
    }
    const float n0 = fmaxf(m0, quad_max(c0)), n1 = fmaxf(m1, quad_max(c1));
    float a0 = 0.f, a1 = 0.f;
    if (n0 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) a0 += __expf(s[nt][0] - n0) + __expf(s[nt][1] - n0);
    }
    if (n1 > -INFINITY) {
#pragma unroll
      for (int nt = 0; nt < 8; nt++) a1 += __expf(s[nt][2] - n1) + __expf(s[nt][3] - n1);
    }
    l0 = (n0 > -INFINITY ? l0 * __expf(m0 - n0) : 0.f) + quad_sum(a0);
    l1 = (n1 > -INFINITY ? l1 * __expf(m1 - n1) : 0.f) + quad_sum(a1);
    m0 = n0; m1 = n1;
  }
  const float inv0 = (r0 < N && l0 > 0.f) ? 1.0f / l0 : 0.f, inv1 = (r1 < N && l1 > 0.f) ? 1.0f / l1 : 0.f;
  if (!(m0 > -INFINITY)) m0 = 0.f;
  if (!(m1 > -INFINITY)) m1 = 0.f;
  // ---------------- pass 2: P = exp(s - m) / l (written once, over the parked scores), O = P V
  float o[DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_split<DH>(sV, base + 2 * D, 3 * D, kc * LQ, N, tid);
    float s[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
      s[nt][0] = (r0 < N && col < N) ? Pb[(long long)r0 * N + col] : -INFINITY;
      s[nt][1] = (r0 < N && col + 1 < N) ? Pb[(long long)r0 * N + col + 1] : -INFINITY;
      s[nt][2] = (r1 < N && col < N) ? Pb[(long long)r1 * N + col] : -INFINITY;
      s[nt][3] = (r1 < N && col + 1 < N) ? Pb[(long long)r1 * N + col + 1] : -INFINITY;
    }
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int col = kc * LQ + nt * 8 + 2 * t;
      s[nt][0] = __expf(s[nt][0] - m0) * inv0; s[nt][1] = __expf(s[nt][1] - m0) * inv0;
      s[nt][2] = __expf(s[nt][2] - m1) * inv1; s[nt][3] = __expf(s[nt][3] - m1) * inv1;
      if (r0 < N) { if (col < N) Pb[(long long)r0 * N + col] = s[nt][0]; if (col + 1 < N) Pb[(long long)r0 * N + col + 1] = s[nt][1]; }   // atg__
      if (r1 < N) { if (col < N) Pb[(long long)r1 * N + col] = s[nt][2]; if (col + 1 < N) Pb[(long long)r1 * N + col + 1] = s[nt][3]; }
    }
    gemm_regA_kn<DH>(o, s, sV, 0, 4, lane);            // P feeds the MMA straight from its accumulator fragments
  }
  // causal: the skipped upper-triangular blocks of P are exactly zero
  if (nchunks < nall) {
    const int w = (nall - nchunks) * LQ;
    for (int e = tid; e < LQ * w; e += 128) {
      const int i = e / w, c = nchunks * LQ + e % w;
      if (q0 + i < N && c < N) Pb[(long long)(q0 + i) * N + c] = 0.f;
    }
  }
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) {
    const int col = h * DH + nd * 8 + 2 * t;
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int row = hf ? r1 : r0;
      if (row < N) {
        const long long oi = ((long long)b * N + row) * D + col;
        float2 v = make_float2(o[nd][2 * hf], o[nd][2 * hf + 1]);
        if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
        *reinterpret_cast<float2*>(out + oi) = v;
        if (o_hi) { uint32_t hh, ll; split2(v.x, v.y, hh, ll); o_hi[oi >> 1] = hh; o_lo[oi >> 1] = ll; }
      }
    }
  }
}

// dP tile (fragment layout) of this warp's 16 query rows against the 64 keys staged in sV, and the matching P tile
template <int DH>
__device__ __forceinline__ void dp_and_p(float (&dp)[8][4], float (&p)[8][4], const SArr& sDO, const SArr& sV, int i0, int r0,
                                         int r1, int kc, int N, const float* __restrict__ Pb, int lane, int t) {
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    const int col = kc * LQ + nt * 8 + 2 * t;
    p[nt][0] = (r0 < N && col < N) ? Pb[(long long)r0 * N + col] : 0.f;
    p[nt][1] = (r0 < N && col + 1 < N) ? Pb[(long long)r0 * N + col + 1] : 0.f;
    p[nt][2] = (r1 < N && col < N) ? Pb[(long long)r1 * N + col] : 0.f;
    p[nt][3] = (r1 < N && col + 1 < N) ? Pb[(long long)r1 * N + col + 1] : 0.f;
    dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
  }
  gemm_rowk_nk<DH>(dp, sDO, i0, sV, 0, 4, lane);                               // dP = dO V^T   attention.py:74
}

// backward, per query block: row sums t (-> tsum) and dQ
template <int DH>
__global__ void __launch_bounds__(128, 3) attn_bwd_bf16_q_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                              const float* __restrict__ P, float* __restrict__ dqkv,
                                                              float* __restrict__ tsum, int N, int H, int causal,
                                                              uint32_t* __restrict__ g_hi, uint32_t* __restrict__ g_lo,
                                                              const float* __restrict__ att_out) {
  pdl_prologue();
  using M = Smem<DH>;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const SArr sDO = make_arr(smem, M::LD), sK = make_arr(smem + M::ARR, M::LD), sV = make_arr(smem + 2 * M::ARR, M::LD);
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int q0 = blockIdx.x * LQ, D = H * DH;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  const float* Pb = P + (long long)bh * N * N;
  const int nall = (N + LQ - 1) / LQ;
  const int nchunks = causal ? min(nall, (int)blockIdx.x + 1) : nall;
  const int i0 = warp * 16, r0 = q0 + i0 + g, r1 = r0 + 8;
  const float scale = rsqrtf((float)DH);
  load_split<DH>(sDO, dout + (long long)b * N * D + h * DH, D, q0, N, tid);
  // pass 1: t_i = sum_j dP_ij P_ij.  With the forward's attention output O = P V at hand this is the row dot product
  // dO_i . O_i (sum_j P_ij (dO_i . V_j) = dO_i . sum_j P_ij V_j), exact fp32 and no sweep over the key blocks.
  float t0 = 0.f, t1 = 0.f;
  if (att_out) {
    const float* dob = dout + (long long)b * N * D + h * DH;
    const float* aob = att_out + (long long)b * N * D + h * DH;
#pragma unroll
    for (int nd = 0; nd < DH / 8; nd++) {
      const int col = nd * 8 + 2 * t;
      if (r0 < N) {
        const float2 a = *reinterpret_cast<const float2*>(dob + (long long)r0 * D + col);
        const float2 o = *reinterpret_cast<const float2*>(aob + (long long)r0 * D + col);
        t0 = fmaf(a.x, o.x, fmaf(a.y, o.y, t0));
      }
      if (r1 < N) {
        const float2 a = *reinterpret_cast<const float2*>(dob + (long long)r1 * D + col);
        const float2 o = *reinterpret_cast<const float2*>(aob + (long long)r1 * D + col);
        t1 = fmaf(a.x, o.x, fmaf(a.y, o.y, t1));
      }
    }
  } else {
    for (int kc = 0; kc < nchunks; kc++) {
      __syncthreads();
      load_split<DH>(sV, base + 2 * D, 3 * D, kc * LQ, N, tid);
      __syncthreads();
      float dp[8][4], p[8][4];
      dp_and_p<DH>(dp, p, sDO, sV, i0, r0, r1, kc, N, Pb, lane, t);
#pragma unroll
      for (int nt = 0; nt < 8; nt++) {
        t0 += dp[nt][0] * p[nt][0] + dp[nt][1] * p[nt][1];
        t1 += dp[nt][2] * p[nt][2] + dp[nt][3] * p[nt][3];
      }
    }
  }
  t0 = quad_sum(t0); t1 = quad_sum(t1);
  if (t == 0) {
    if (r0 < N) tsum[(long long)bh * N + r0] = t0;
    if (r1 < N) tsum[(long long)bh * N + r1] = t1;
  }
  // pass 2: dS = P o (dP - t) (attention.py:75), dQ += dS K
  float dq[DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;
  for (int kc = 0; kc < nchunks; kc++) {
    __syncthreads();
    load_split<DH>(sV, base + 2 * D, 3 * D, kc * LQ, N, tid);
    load_split<DH>(sK, base + D, 3 * D, kc * LQ, N, tid);
    __syncthreads();
    float dp[8][4], p[8][4];
    dp_and_p<DH>(dp, p, sDO, sV, i0, r0, r1, kc, N, Pb, lane, t);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      dp[nt][0] = p[nt][0] * (dp[nt][0] - t0); dp[nt][1] = p[nt][1] * (dp[nt][1] - t0);
      dp[nt][2] = p[nt][2] * (dp[nt][2] - t1); dp[nt][3] = p[nt][3] * (dp[nt][3] - t1);
    }
    gemm_regA_kn<DH>(dq, dp, sK, 0, 4, lane);                                  // dQ = dS K    attention.py:78
  }
  float* ob = dqkv + (long long)b * N * 3 * D + h * DH;
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) {
    const int col = nd * 8 + 2 * t;
    if (r0 < N) store_g(ob, g_hi, g_lo, dqkv, (long long)r0 * 3 * D + col, dq[nd][0] * scale, dq[nd][1] * scale);
    if (r1 < N) store_g(ob, g_hi, g_lo, dqkv, (long long)r1 * 3 * D + col, dq[nd][2] * scale, dq[nd][3] * scale);
  }
}

// backward, per key block: dK = dS^T Q / sqrt(dh), dV = P^T dO, with dS recomputed from dO, V, P and the row sums
template <int DH>
__global__ void __launch_bounds__(128) attn_bwd_bf16_k_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                              const float* __restrict__ P, float* __restrict__ dqkv,
                                                              const float* __restrict__ tsum, int N, int H, int causal,
                                                              uint32_t* __restrict__ g_hi, uint32_t* __restrict__ g_lo) {
  pdl_prologue();
  using M = Smem<DH>;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const SArr sV = make_arr(smem, M::LD), sQ = make_arr(smem + M::ARR, M::LD), sDO = make_arr(smem + 2 * M::ARR, M::LD);
  const SArr sP = make_arr(smem + 3 * M::ARR, LDP), sdS = make_arr(smem + 3 * M::ARR + M::PARR, LDP);   // [64 queries][64 keys]
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  const int k0 = blockIdx.x * LQ, D = H * DH;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  const float* Pb = P + (long long)bh * N * N;
  const int nall = (N + LQ - 1) / LQ;
  const int qfirst = causal ? (int)blockIdx.x : 0;     // query blocks before the key block see only zeros
  const int i0 = warp * 16;
  const float scale = rsqrtf((float)DH);
  load_split<DH>(sV, base + 2 * D, 3 * D, k0, N, tid);
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++)
#pragma unroll
    for (int c = 0; c < 4; c++) dk[nd][c] = dv[nd][c] = 0.f;
  for (int qc = qfirst; qc < nall; qc++) {
    __syncthreads();
    load_split<DH>(sQ, base, 3 * D, qc * LQ, N, tid);
    load_split<DH>(sDO, dout + (long long)b * N * D + h * DH, D, qc * LQ, N, tid);
    __syncthreads();
    // this warp's 16 QUERY rows of the (query block x key block) tile: dP, P, dS -> shared (split)
    const int r0 = qc * LQ + i0 + g, r1 = r0 + 8;
    float dp[8][4], p[8][4];
    dp_and_p<DH>(dp, p, sDO, sV, i0, r0, r1, (int)blockIdx.x, N, Pb, lane, t);
    const float t0 = r0 < N ? tsum[(long long)bh * N + r0] : 0.f, t1 = r1 < N ? tsum[(long long)bh * N + r1] : 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int c = nt * 8 + 2 * t;
      st2(sP, i0 + g, c, p[nt][0], p[nt][1]);
      st2(sP, i0 + g + 8, c, p[nt][2], p[nt][3]);
      st2(sdS, i0 + g, c, p[nt][0] * (dp[nt][0] - t0), p[nt][1] * (dp[nt][1] - t0));
      st2(sdS, i0 + g + 8, c, p[nt][2] * (dp[nt][2] - t1), p[nt][3] * (dp[nt][3] - t1));
    }
    __syncthreads();
    // now the warp owns 16 KEYS (columns i0.. of the tile), contraction over the 64 queries of the block
    gemm_colk_kn<DH>(dk, sdS, i0, sQ, 0, 4, lane);                             // dK = dS^T Q   attention.py:79
    gemm_colk_kn<DH>(dv, sP, i0, sDO, 0, 4, lane);                             // dV = P^T dO   attention.py:76
  }
  float* ob = dqkv + (long long)b * N * 3 * D + h * DH;
  const int j0 = k0 + i0 + g, j1 = j0 + 8;
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) {
    const int col = nd * 8 + 2 * t;
    if (j0 < N) {
      store_g(ob, g_hi, g_lo, dqkv, (long long)j0 * 3 * D + col + D, dk[nd][0] * scale, dk[nd][1] * scale);
      store_g(ob, g_hi, g_lo, dqkv, (long long)j0 * 3 * D + col + 2 * D, dv[nd][0], dv[nd][1]);
    }
    if (j1 < N) {
      store_g(ob, g_hi, g_lo, dqkv, (long long)j1 * 3 * D + col + D, dk[nd][2] * scale, dk[nd][3] * scale);
      store_g(ob, g_hi, g_lo, dqkv, (long long)j1 * 3 * D + col + 2 * D, dv[nd][2], dv[nd][3]);
    }
  }
}

// ---------------------------------------------------------------- single-block kernels (N <= 64: the scaled ViT, N = 49)
// One CTA per (sample, head), warp w owns query rows 16w..: scores, softmax and P V in one pass over registers.
template <int DH>
__global__ void __launch_bounds__(128) attn_fwd_bf16_small_kernel(const float* __restrict__ qkv, const float* __restrict__ mask,
                                                                  int mask_kind, float* __restrict__ out, float* __restrict__ P,
                                                                  float* __restrict__ S, int N, int H,
                                                                  const float* __restrict__ residual,
                                                                  uint32_t* __restrict__ o_hi, uint32_t* __restrict__ o_lo) {
  pdl_prologue();
  using M = Smem<DH>;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const SArr sQ = make_arr(smem, M::LD), sK = make_arr(smem + M::ARR, M::LD), sV = make_arr(smem + 2 * M::ARR, M::LD);
  const int bh = blockIdx.x, b = bh / H, h = bh % H, D = H * DH;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  load_split<DH>(sQ, base, 3 * D, 0, N, tid);
  load_split<DH>(sK, base + D, 3 * D, 0, N, tid);
  load_split<DH>(sV, base + 2 * D, 3 * D, 0, N, tid);
  __syncthreads();
  const int i0 = warp * 16, r0 = i0 + g, r1 = r0 + 8;
  if (i0 >= N) return;                       // no CTA-wide barrier below
  float* Pb = P + (long long)bh * N * N;
  float s[8][4];
  scores<DH>(s, sQ, sK, i0, r0, r1, 0, N, rsqrtf((float)DH), mask_kind, mask, S ? S + (long long)bh * N * N : nullptr, lane, t);
  float c0 = -INFINITY, c1 = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) { c0 = fmaxf(c0, fmaxf(s[nt][0], s[nt][1])); c1 = fmaxf(c1, fmaxf(s[nt][2], s[nt][3])); }
  float m0 = quad_max(c0), m1 = quad_max(c1);
  if (!(m0 > -INFINITY)) m0 = 0.f;
  if (!(m1 > -INFINITY)) m1 = 0.f;
  float a0 = 0.f, a1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    s[nt][0] = __expf(s[nt][0] - m0); s[nt][1] = __expf(s[nt][1] - m0);
    s[nt][2] = __expf(s[nt][2] - m1); s[nt][3] = __expf(s[nt][3] - m1);
    a0 += s[nt][0] + s[nt][1];
    a1 += s[nt][2] + s[nt][3];
  }
  a0 = quad_sum(a0); a1 = quad_sum(a1);
  const float inv0 = (r0 < N && a0 > 0.f) ? 1.0f / a0 : 0.f, inv1 = (r1 < N && a1 > 0.f) ? 1.0f / a1 : 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    const int col = nt * 8 + 2 * t;
    s[nt][0] *= inv0; s[nt][1] *= inv0; s[nt][2] *= inv1; s[nt][3] *= inv1;
    if (r0 < N) { if (col < N) Pb[r0 * N + col] = s[nt][0]; if (col + 1 < N) Pb[r0 * N + col + 1] = s[nt][1]; }      // atg__
    if (r1 < N) { if (col < N) Pb[r1 * N + col] = s[nt][2]; if (col + 1 < N) Pb[r1 * N + col + 1] = s[nt][3]; }
  }
  float o[DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) o[nd][0] = o[nd][1] = o[nd][2] = o[nd][3] = 0.f;
  gemm_regA_kn<DH>(o, s, sV, 0, 4, lane);              // P feeds the MMA straight from its accumulator fragments
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) {
    const int col = h * DH + nd * 8 + 2 * t;
#pragma unroll
    for (int hf = 0; hf < 2; hf++) {
      const int row = hf ? r1 : r0;
      if (row < N) {
        const long long oi = ((long long)b * N + row) * D + col;
        float2 v = make_float2(o[nd][2 * hf], o[nd][2 * hf + 1]);
        if (residual) { const float2 rr = __ldg(reinterpret_cast<const float2*>(residual + oi)); v.x += rr.x; v.y += rr.y; }
        *reinterpret_cast<float2*>(out + oi) = v;
        if (o_hi) { uint32_t hh, ll; split2(v.x, v.y, hh, ll); o_hi[oi >> 1] = hh; o_lo[oi >> 1] = ll; }
      }
    }
  }
}

template <int DH>
__global__ void __launch_bounds__(128, 3) attn_bwd_bf16_small_kernel(const float* __restrict__ dout, const float* __restrict__ qkv,
                                                                  const float* __restrict__ P, float* __restrict__ dqkv, int N,
                                                                  int H, uint32_t* __restrict__ g_hi, uint32_t* __restrict__ g_lo) {
  pdl_prologue();
  using M = Smem<DH>;
  extern __shared__ __align__(128) char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const SArr sQ = make_arr(smem, M::LD), sK = make_arr(smem + M::ARR, M::LD), sV = make_arr(smem + 2 * M::ARR, M::LD);
  const SArr sDO = make_arr(smem + 3 * M::ARR, M::LD);
  static_assert(M::PARR <= M::ARR, "P / dS reuse the K / V buffers");
  const SArr sP = make_arr(smem + M::ARR, LDP), sdS = make_arr(smem + 2 * M::ARR, LDP);     // alias sK / sV (phase 2 only)
  const int bh = blockIdx.x, b = bh / H, h = bh % H, D = H * DH;
  const float* base = qkv + (long long)b * N * 3 * D + h * DH;
  const float* Pb = P + (long long)bh * N * N;
  load_split<DH>(sQ, base, 3 * D, 0, N, tid);
  load_split<DH>(sK, base + D, 3 * D, 0, N, tid);
  load_split<DH>(sV, base + 2 * D, 3 * D, 0, N, tid);
  load_split<DH>(sDO, dout + (long long)b * N * D + h * DH, D, 0, N, tid);
  __syncthreads();
  const int i0 = warp * 16, r0 = i0 + g, r1 = r0 + 8;
  const float scale = rsqrtf((float)DH);
  float* ob = dqkv + (long long)b * N * 3 * D + h * DH;
  {
    // phase 1 (this warp's 16 QUERIES): dP, dS = P o (dP - rowsum(dP o P)), dQ = dS K / sqrt(dh)   attention.py:74-78
    float dp[8][4], p[8][4];
    dp_and_p<DH>(dp, p, sDO, sV, i0, r0, r1, 0, N, Pb, lane, t);
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      t0 += dp[nt][0] * p[nt][0] + dp[nt][1] * p[nt][1];
      t1 += dp[nt][2] * p[nt][2] + dp[nt][3] * p[nt][3];
    }
    t0 = quad_sum(t0); t1 = quad_sum(t1);
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      dp[nt][0] = p[nt][0] * (dp[nt][0] - t0); dp[nt][1] = p[nt][1] * (dp[nt][1] - t0);
      dp[nt][2] = p[nt][2] * (dp[nt][2] - t1); dp[nt][3] = p[nt][3] * (dp[nt][3] - t1);
    }
    float dq[DH / 8][4];
#pragma unroll
    for (int nd = 0; nd < DH / 8; nd++) dq[nd][0] = dq[nd][1] = dq[nd][2] = dq[nd][3] = 0.f;
    gemm_regA_kn<DH>(dq, dp, sK, 0, 4, lane);            // dS feeds the MMA straight from its accumulator fragments
#pragma unroll
    for (int nd = 0; nd < DH / 8; nd++) {
      const int col = nd * 8 + 2 * t;
      if (r0 < N) store_g(ob, g_hi, g_lo, dqkv, (long long)r0 * 3 * D + col, dq[nd][0] * scale, dq[nd][1] * scale);
      if (r1 < N) store_g(ob, g_hi, g_lo, dqkv, (long long)r1 * 3 * D + col, dq[nd][2] * scale, dq[nd][3] * scale);
    }
    // K and V are dead from here on: P and dS (needed TRANSPOSED by phase 2) take their place in shared memory
    __syncthreads();
#pragma unroll
    for (int nt = 0; nt < 8; nt++) {
      const int c = nt * 8 + 2 * t;
      st2(sP, r0, c, p[nt][0], p[nt][1]);
      st2(sP, r1, c, p[nt][2], p[nt][3]);
      st2(sdS, r0, c, dp[nt][0], dp[nt][1]);
      st2(sdS, r1, c, dp[nt][2], dp[nt][3]);
    }
  }
  __syncthreads();
  if (i0 >= N) return;
  // phase 2 (this warp's 16 KEYS): dK = dS^T Q / sqrt(dh), dV = P^T dO                               attention.py:76,79
  float dk[DH / 8][4], dv[DH / 8][4];
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++)
#pragma unroll
    for (int c = 0; c < 4; c++) dk[nd][c] = dv[nd][c] = 0.f;
  gemm_colk_kn<DH>(dk, sdS, i0, sQ, 0, 4, lane);
  gemm_colk_kn<DH>(dv, sP, i0, sDO, 0, 4, lane);
#pragma unroll
  for (int nd = 0; nd < DH / 8; nd++) {
    const int col = nd * 8 + 2 * t;
    if (r0 < N) {
      store_g(ob, g_hi, g_lo, dqkv, (long long)r0 * 3 * D + col + D, dk[nd][0] * scale, dk[nd][1] * scale);
      store_g(ob, g_hi, g_lo, dqkv, (long long)r0 * 3 * D + col + 2 * D, dv[nd][0], dv[nd][1]);
    }
    if (r1 < N) {
      store_g(ob, g_hi, g_lo, dqkv, (long long)r1 * 3 * D + col + D, dk[nd][2] * scale, dk[nd][3] * scale);
      store_g(ob, g_hi, g_lo, dqkv, (long long)r1 * 3 * D + col + 2 * D, dv[nd][2], dv[nd][3]);
    }
  }
}

template <int DH> static size_t fwd_smem() { return 3 * (size_t)Smem<DH>::ARR; }
template <int DH> static size_t bwdq_smem() { return 3 * (size_t)Smem<DH>::ARR; }
template <int DH> static size_t bwdk_smem() { return 3 * (size_t)Smem<DH>::ARR + 2 * (size_t)Smem<DH>::PARR; }
template <int DH> static size_t bwds_smem() { return 4 * (size_t)Smem<DH>::ARR; }

}  // namespace ab

// dh == 64 (scaled ViT N = 49 -> single-block kernels; GPT configs T = 256 -> two-pass kernels); NTB200_ATT_BF16=0
// keeps the TF32x3 kernels of attention.cu, NTB200_ATT_BF16_MIN_N (default 16) moves the sequence-length threshold.
// Measured on B200 per step: G1 forward 1.03 -> 0.76 ms, backward 1.50 -> 1.09 ms; V2 forward 2.80 -> 2.54 ms,
// backward 3.66 -> 2.49 ms.
bool attention_bf16_ok(int N, int H, int dh) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NTB200_ATT_BF16"); on = (e && e[0] == '0') ? 0 : 1; }
  const char* f = getenv("NTB200_ATT_BF16_MIN_N");
  const int min_n = f ? atoi(f) : 16;         // below ~16 tokens the TF32x3 single-block kernels are faster (G0, T = 5: -7 us/step)
  // mode 2 (bar 1e-2) keeps these kernels: their 3-pass products are more accurate than it asks for and faster than TF32x3
  return on && dh == 64 && N >= min_n && g_gemm_mode >= 1;
}

// opt-in to > 48 KB of dynamic shared memory, once (the backward may run without this library's forward having run: the
// tcgen05 forward of attention_tcl.cu feeds it)
static int configure() {
  using namespace ab;
  static bool cfg = false;
  if (cfg) return 0;
  NT_CHECK(cudaFuncSetAttribute(attn_fwd_bf16_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<64>()));
  NT_CHECK(cudaFuncSetAttribute(attn_bwd_bf16_q_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwdq_smem<64>()));
  NT_CHECK(cudaFuncSetAttribute(attn_bwd_bf16_k_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwdk_smem<64>()));
  NT_CHECK(cudaFuncSetAttribute(attn_fwd_bf16_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fwd_smem<64>()));
  NT_CHECK(cudaFuncSetAttribute(attn_bwd_bf16_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwds_smem<64>()));
  cfg = true;
  return 0;
}

int attention_fwd_bf16(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N, int H,
                       const float* residual, void* o_hi, void* o_lo) {
  using namespace ab;
  if (configure()) return 1;
  if (N <= LQ) {
    NT_CHECK(launch_k(attn_fwd_bf16_small_kernel<64>, dim3(B * H), dim3(128), fwd_smem<64>(), qkv, mask, mask_kind, out, P, S, N, H,
                      residual, (uint32_t*)o_hi, (uint32_t*)o_lo));
    NT_LAUNCH_OK();
    return 0;
  }
  NT_CHECK(launch_k(attn_fwd_bf16_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), fwd_smem<64>(), qkv, mask, mask_kind, out, P,
                    S, N, H, residual, (uint32_t*)o_hi, (uint32_t*)o_lo));
  NT_LAUNCH_OK();
  return 0;
}

int attention_bwd_bf16(const float* dout, const float* qkv, const float* P, float* dqkv, float* tsum, int B, int N, int H,
                       int causal, void* g_hi, void* g_lo, const float* att_out) {
  using namespace ab;
  if (configure()) return 1;
  if (N <= LQ) {
    NT_CHECK(launch_k(attn_bwd_bf16_small_kernel<64>, dim3(B * H), dim3(128), bwds_smem<64>(), dout, qkv, P, dqkv, N, H, (uint32_t*)g_hi, (uint32_t*)g_lo));
    NT_LAUNCH_OK();
    return 0;
  }
  NT_CHECK(launch_k(attn_bwd_bf16_q_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), bwdq_smem<64>(), dout, qkv, P, dqkv, tsum,
                    N, H, causal, (uint32_t*)g_hi, (uint32_t*)g_lo, att_out));
  NT_LAUNCH_OK();
  NT_CHECK(launch_k(attn_bwd_bf16_k_kernel<64>, dim3(ceil_div(N, LQ), B * H), dim3(128), bwdk_smem<64>(), dout, qkv, P, dqkv,
                    (const float*)tsum, N, H, causal, (uint32_t*)g_hi, (uint32_t*)g_lo));
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
