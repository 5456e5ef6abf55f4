// Fused ViT encoder blocks on the 5th-generation tensor cores (attention.py:20-89), the launch-bound README shape
// (N <= 64 tokens, D = 96, 4 heads of 24, MLP 2x): ONE launch runs every block of the stack for every image, forward or
// backward.  One CTA per image; every contraction is a tcgen05.mma (kind::f16, bf16 operands, fp32 accumulator in TMEM)
// issued by a single thread, error-compensated (x = hi + lo in bf16, product = lo*hi + hi*lo + hi*hi: fp32-class, ~1e-5).
//
//   * Operands live in shared memory as bf16 hi / lo PLANES in the canonical 128-byte-swizzled layout (umma.cuh): an
//     activation [token][feature] is written once by the epilogue threads and then read by the tensor core as a K-major
//     operand (Y = X W), and — the same bytes — as an MN-major operand (dW = X^T dY, P V, dS^T Q): nothing is transposed.
//   * M = 128 instructions with the image's 49 tokens in rows 0..48: accumulator row i is TMEM lane i, so one THREAD owns
//     one token row and LayerNorm / softmax / GELU / the hi-lo split are thread-local (no shuffles); rows 64..127 of the
//     A operand alias whatever follows in shared memory — their accumulator rows are never read.
//   * Heads are padded from 24 to 32 columns in the q / k / v planes so that every head starts on a K=16 step; a head is a
//     64-byte start offset inside the swizzled row.
//   * Weights are re-laid once per step (vit_tc_wprep_kernel) into ready-to-use operand tiles of [96 out][64 in] hi + lo
//     (24 KB) and streamed through a ring by 1-D bulk copies (TMA engine) issued by one thread, running ahead of the MMAs.
//   * Warp roles: 8 epilogue warps (TMEM lane quadrant x column group; column group g of the residual stream IS head g),
//     1 loader thread, 1 MMA-issuing thread; all hand-offs are mbarriers (tcgen05.commit on the MMA side).
//   * The fp32 residual stream of the image stays in REGISTERS (24 values per thread) across the whole stack.
#include "common.cuh"
#include "umma.cuh"

namespace nt {
namespace vtc {
using namespace um;

constexpr int MAXB = 16;      // blocks per launch (kernel parameter space)
struct Args {
  NtVitBlock blk[MAXB];
  int nblocks;
  int N;
  const float* x;       // [B,N,D] input of block 0
  const float* dout;    // [B,N,D] gradient w.r.t. the last block's output      (backward)
  float* dx;            // [B,N,D] gradient w.r.t. x                              (backward)
};

constexpr int D = 96, H = 4, DH = 24, DHP = 32, HD = 192;
constexpr int ROWS = 64;                       // token rows of a plane chunk
constexpr uint32_t CH = ROWS * 128;            // bytes of one 64-column chunk of a plane (8 KB)
constexpr int NT = 96;                         // rows (output features) of a weight tile
constexpr uint32_t WPLANE = NT * 128;          // 12 KB
constexpr uint32_t STAGE = 2 * WPLANE;         // hi + lo
constexpr int FWD_STAGES = 13, BWD_STAGES = 13, ALL_STAGES = FWD_STAGES + BWD_STAGES;
constexpr int NSTAGE = 4;
constexpr int THREADS = 448;                   // 14 warps: epilogue {0,1,4,5,8,9,12,13}, loader 2, MMA 3
constexpr int EPI_THREADS = 256;

// shared memory map (bytes)
constexpr uint32_t OFF_A = 0;                        // 4 chunks: U / U1 (hi 2, lo 2)  |  P0 hi, P0 lo, P1 hi, P1 lo
constexpr uint32_t OFF_B = OFF_A + 4 * CH;           // 12 chunks: Q hi 2, Q lo 2, K hi 2, K lo 2, V hi 2, V lo 2  |  G hi 3, G lo 3
constexpr uint32_t OFF_RING = OFF_B + 12 * CH;
constexpr uint32_t OFF_STAT = OFF_RING + NSTAGE * STAGE;      // [2][64][4] floats
constexpr uint32_t OFF_BAR = OFF_STAT + 2 * 64 * 4 * 4;
constexpr int NBAR = 2 * NSTAGE + 24;
constexpr uint32_t SMEM_BYTES = OFF_BAR + NBAR * 8 + 16;          // no static shared memory: the dynamic window starts 1024-aligned
static_assert(SMEM_BYTES <= 232448, "shared memory budget");

// ---------------------------------------------------------------- weight images
// W [K][Nw] row-major fp32 -> operand tiles.  Stage order per block:
//   forward  0..5  qkv  (tile nt = i/2 in {q,k,v}, k-chunk i%2)        B[n][k] = Wqkv[64kc+k][96nt+n]
//            6..9  fc0  (tile nt, k-chunk)                             B[n][k] = W0[64kc+k][96nt+n]
//            10..12 fc1 (k-chunk)                                      B[n][k] = W1[64kc+k][n]
//   backward 13..16 dG = d W1^T   (tile nt, k-chunk)                   B[n][k] = W1[96nt+n][64kc+k]
//            17..19 dU1 = dpre W0^T (k-chunk)                          B[n][k] = W0[n][64kc+k]
//            20..25 dU = dqkv Wqkv^T, k runs over the HEAD-PADDED columns [q|k|v][head][32]
__device__ __forceinline__ float wimg_value(const NtVitBlock& P, int i, int n, int kk) {
  if (i < 6) {
    const int nt = i >> 1, k = 64 * (i & 1) + kk;
    return k < D ? P.wqkv[(size_t)k * 3 * D + NT * nt + n] : 0.f;
  }
  if (i < 10) {
    const int j = i - 6, nt = j >> 1, k = 64 * (j & 1) + kk;
    return k < D ? P.w0[(size_t)k * HD + NT * nt + n] : 0.f;
  }
  if (i < 13) return P.w1[(size_t)(64 * (i - 10) + kk) * D + n];
  i -= 13;
  if (i < 4) {
    const int nt = i >> 1, k = 64 * (i & 1) + kk;
    return k < D ? P.w1[(size_t)(NT * nt + n) * D + k] : 0.f;
  }
  if (i < 7) return P.w0[(size_t)n * HD + 64 * (i - 4) + kk];
  const int kp = 64 * (i - 7) + kk, part = kp >> 7, hh = (kp & 127) >> 5, d = kp & 31;
  return d < DH ? P.wqkv[(size_t)n * 3 * D + part * D + DH * hh + d] : 0.f;
}

__global__ void __launch_bounds__(256) vit_tc_wprep_kernel(Args a) {
  pdl_prologue();
  const NtVitBlock& P = a.blk[blockIdx.y];
  uint8_t* img = reinterpret_cast<uint8_t*>(P.wfrag);
  const int units = ALL_STAGES * NT * 8;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < units; e += gridDim.x * blockDim.x) {
    const int u = e & 7, n = (e >> 3) % NT, i = e / (8 * NT);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; j++) v[j] = wimg_value(P, i, n, 8 * u + j);
    uint4 h, l;
    split2(v[0], v[1], h.x, l.x);
    split2(v[2], v[3], h.y, l.y);
    split2(v[4], v[5], h.z, l.z);
    split2(v[6], v[7], h.w, l.w);
    const uint32_t off = (uint32_t)i * STAGE + (uint32_t)n * 128u + (((uint32_t)u ^ ((uint32_t)n & 7u)) << 4);
    *reinterpret_cast<uint4*>(img + off) = h;
    *reinterpret_cast<uint4*>(img + off + WPLANE) = l;
  }
}

// ---------------------------------------------------------------- pieces shared by forward and backward
struct Bars {
  uint64_t* full;     // [NSTAGE]
  uint64_t* empty;    // [NSTAGE]
  uint64_t* b;        // role-specific barriers
};

// exact-erf GELU and its derivative from one shared exponential (net/activation.py:139-148).
// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): erf(s) = 1 - (a1 t + .. + a5 t^5) exp(-s^2), t = 1/(1 + p|s|).
__device__ __forceinline__ void gelu_both(float x, float& y, float& dy) {
  const float s = x * 0.70710678118654752f;
  const float e = __expf(-s * s);
  const float t = __fdividef(1.0f, fmaf(0.3275911f, fabsf(s), 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float er = copysignf(fmaf(-poly * t, e, 1.0f), s);
  const float cdf = fmaf(0.5f, er, 0.5f);
  y = x * cdf;
  dy = fmaf(x * 0.3989422804014327f, e, cdf);
}

// the ring consumer side: issue the MMAs of one weight stage.  A planes at (a_hi, a_lo) (byte addresses of the chunk the
// stage contracts over), `ksteps` K=16 steps, accumulator `tacc`; `first` = the accumulator starts from zero.
__device__ __forceinline__ void mma_stage(uint32_t a_hi, uint32_t a_lo, uint32_t stage_addr, int ksteps, uint32_t tacc,
                                          uint32_t idesc, bool first) {
#pragma unroll 1
  for (int ks = 0; ks < ksteps; ks++) {
    const uint64_t ah = desc_kmajor(a_hi + ks * KMAJ_STEP), al = desc_kmajor(a_lo + ks * KMAJ_STEP);
    const uint64_t bh = desc_kmajor(stage_addr + ks * KMAJ_STEP), bl = desc_kmajor(stage_addr + WPLANE + ks * KMAJ_STEP);
    mma_bf16(tacc, al, bh, idesc, (first && ks == 0) ? 0u : 1u);      // small terms first
    mma_bf16(tacc, ah, bl, idesc, 1u);
    mma_bf16(tacc, ah, bh, idesc, 1u);
  }
}

// one thread's share (24 columns) of a row-wise mean / rstd over the D = 96 columns held by the 4 column groups
__device__ __forceinline__ void row_stats(const float (&v)[24], float* stat, int r, int g, float& mean, float& rstd) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < 24; c++) s += v[c];
  stat[r * 4 + g] = s;
  named_bar_sync(1, EPI_THREADS);
  const float4 t = *reinterpret_cast<const float4*>(stat + r * 4);
  mean = (t.x + t.y + t.z + t.w) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < 24; c++) { const float d = v[c] - mean; q += d * d; }
  stat[256 + r * 4 + g] = q;
  named_bar_sync(1, EPI_THREADS);
  const float4 u = *reinterpret_cast<const float4*>(stat + 256 + r * 4);
  rstd = rsqrtf((u.x + u.y + u.z + u.w) * (1.0f / D) + 1e-5f);       // biased variance, eps inside (layernorm.py:100-103)
}

__device__ __forceinline__ void ld24(uint32_t taddr, float (&v)[24]) {
  uint32_t a[16], b[8];
  tmem_ld16(taddr, a);
  tmem_ld8(taddr + 16, b);
  tmem_wait_ld();
#pragma unroll
  for (int j = 0; j < 16; j++) v[j] = __uint_as_float(a[j]);
#pragma unroll
  for (int j = 0; j < 8; j++) v[16 + j] = __uint_as_float(b[j]);
}

// 24 consecutive columns of row r into a hi/lo plane pair starting at column c0 (c0 % 8 == 0)
__device__ __forceinline__ void store24_split(uint8_t* hi, uint8_t* lo, int r, int c0, const float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 3; j++) {
    float w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = v[8 * j + i];
    store8_split(hi, lo, r, c0 + 8 * j, CH, w);
  }
}
__device__ __forceinline__ void store24_global(float* p, const float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 6; j++) *reinterpret_cast<float4*>(p + 4 * j) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void load24_global(const float* p, float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const float4 t = *reinterpret_cast<const float4*>(p + 4 * j);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
}

// barrier indices (forward)
enum { FB_U = 0, FB_QKVACC = 1 /*3*/, FB_QKVRDY = 4, FB_S = 5 /*4*/, FB_P = 9 /*4*/, FB_O = 13 /*4*/, FB_U1 = 17, FB_F0 = 18 /*2*/,
       FB_G = 20, FB_F1 = 21, FB_COUNT = 22 };

// ---------------------------------------------------------------- forward
__global__ void __launch_bounds__(THREADS, 1) vit_tc_fwd_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();               // SWIZZLE_128B operands need 1024-byte aligned chunks
  uint8_t* sA = smem + OFF_A;
  uint8_t* sB = smem + OFF_B;
  uint8_t* sRing = smem + OFF_RING;
  float* stat = reinterpret_cast<float*>(smem + OFF_STAT);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* empty = full + NSTAGE;
  uint64_t* bar = empty + NSTAGE;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 24);

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, img = blockIdx.x;
  const bool is_epi = (warp & 3) < 2 && warp < 14;
  const int q = warp & 3, g = warp >> 2, r = 32 * q + lane;      // epilogue: TMEM lane quadrant, column group (= head), token row

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(&bar[FB_U], 8);
    for (int i = 0; i < 3; i++) mbar_init(&bar[FB_QKVACC + i], 1);
    mbar_init(&bar[FB_QKVRDY], 8);
    for (int i = 0; i < 4; i++) { mbar_init(&bar[FB_S + i], 1); mbar_init(&bar[FB_P + i], 2); mbar_init(&bar[FB_O + i], 1); }
    mbar_init(&bar[FB_U1], 8);
    mbar_init(&bar[FB_F0], 1); mbar_init(&bar[FB_F0 + 1], 1);
    mbar_init(&bar[FB_G], 8);
    mbar_init(&bar[FB_F1], 1);
    mbar_fence_init();
  }
  // zero the operand regions once: head-padding columns and padded token rows must be exact zeros / finite forever
  for (uint32_t i = tid * 16; i < 16 * CH; i += THREADS * 16) *reinterpret_cast<uint4*>(smem + OFF_A + i) = make_uint4(0, 0, 0, 0);
  if (warp == 3) tmem_alloc<512>(tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  pdl_wait();

  if (warp == 2) {
    // ===================== loader: weight tiles through the ring, running ahead of the MMAs =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int bi = 0; bi < a.nblocks; bi++) {
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.blk[bi].wfrag);
        for (int i = 0; i < FWD_STAGES; i++, it++) {
          const int s = it % NSTAGE;
          if (it >= NSTAGE) mbar_wait(&empty[s], ((it / NSTAGE) - 1) & 1);
          mbar_expect_tx(&full[s], STAGE);
          bulk_g2s(sRing + s * STAGE, wimg + (size_t)i * STAGE, STAGE, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t aA = smem_u32(sA), aB = smem_u32(sB), aR = smem_u32(sRing);
      const uint32_t id_lin = idesc_bf16(128, NT, 0, 0), id_s = idesc_bf16(128, 64, 0, 0), id_pv = idesc_bf16(128, DHP, 0, 1);
      const uint32_t Qh = aB, Ql = aB + 2 * CH, Kh = aB + 4 * CH, Kl = aB + 6 * CH, Vh = aB + 8 * CH, Vl = aB + 10 * CH;
      uint32_t it = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % NSTAGE;
        mbar_wait(&full[s], (it / NSTAGE) & 1);
        tc_fence_after();
        return aR + s * STAGE;
      };
      auto stage_done = [&]() { mma_commit(&empty[it % NSTAGE]); it++; };
      for (int bi = 0; bi < a.nblocks; bi++) {
        const uint32_t ph = bi & 1;
        // ---- qkv = LN(x) Wqkv                                          attention.py:23-24
        mbar_wait(&bar[FB_U], ph);
        tc_fence_after();
        for (int nt = 0; nt < 3; nt++) {
          for (int kc = 0; kc < 2; kc++) {
            const uint32_t st = stage_wait();
            mma_stage(aA + kc * CH, aA + 2 * CH + kc * CH, st, kc ? 2 : 4, tbase + NT * nt, id_lin, kc == 0);
            stage_done();
          }
          mma_commit(&bar[FB_QKVACC + nt]);
        }
        // ---- S_h = Q_h K_h^T                                           attention.py:35-36
        mbar_wait(&bar[FB_QKVRDY], ph);
        tc_fence_after();
        for (int h = 0; h < H; h++) {
          const uint32_t off = (h >> 1) * CH + (h & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {
            const uint64_t qh = desc_kmajor(Qh + off + ks * KMAJ_STEP), ql = desc_kmajor(Ql + off + ks * KMAJ_STEP);
            const uint64_t kh = desc_kmajor(Kh + off + ks * KMAJ_STEP), kl = desc_kmajor(Kl + off + ks * KMAJ_STEP);
            mma_bf16(tbase + 64 * h, ql, kh, id_s, ks ? 1u : 0u);
            mma_bf16(tbase + 64 * h, qh, kl, id_s, 1u);
            mma_bf16(tbase + 64 * h, qh, kh, id_s, 1u);
          }
          mma_commit(&bar[FB_S + h]);
        }
        // ---- O_h = P_h V_h                                             attention.py:42-43
        for (int h = 0; h < H; h++) {
          mbar_wait(&bar[FB_P + h], ph);
          tc_fence_after();
          const uint32_t Ph = (h < 2 ? aA : aB) + (h & 1) * 2 * CH, Pl = Ph + CH;
          const uint32_t off = (h >> 1) * CH + (h & 1) * 64;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {
            const uint64_t p_h = desc_kmajor(Ph + ks * KMAJ_STEP), p_l = desc_kmajor(Pl + ks * KMAJ_STEP);
            const uint64_t vh = desc_mnmajor(Vh + off + ks * MNMAJ_STEP, CH), vl = desc_mnmajor(Vl + off + ks * MNMAJ_STEP, CH);
            mma_bf16(tbase + 288 + DHP * h, p_l, vh, id_pv, ks ? 1u : 0u);
            mma_bf16(tbase + 288 + DHP * h, p_h, vl, id_pv, 1u);
            mma_bf16(tbase + 288 + DHP * h, p_h, vh, id_pv, 1u);
          }
          mma_commit(&bar[FB_O + h]);
        }
        // ---- pre = LN1(h) W0                                           attention.py:49
        mbar_wait(&bar[FB_U1], ph);
        tc_fence_after();
        for (int nt = 0; nt < 2; nt++) {
          for (int kc = 0; kc < 2; kc++) {
            const uint32_t st = stage_wait();
            mma_stage(aA + kc * CH, aA + 2 * CH + kc * CH, st, kc ? 2 : 4, tbase + NT * nt, id_lin, kc == 0);
            stage_done();
          }
          mma_commit(&bar[FB_F0 + nt]);
        }
        // ---- y = GELU(pre) W1                                          attention.py:51
        mbar_wait(&bar[FB_G], ph);
        tc_fence_after();
        for (int kc = 0; kc < 3; kc++) {
          const uint32_t st = stage_wait();
          mma_stage(aB + kc * CH, aB + 3 * CH + kc * CH, st, 4, tbase + 416, id_lin, kc == 0);
          stage_done();
        }
        mma_commit(&bar[FB_F1]);
      }
    }
    __syncwarp();
  } else if (is_epi) {
    // ===================== epilogue warps: one token row per thread, 24 columns (= one head) per column group ========
    const bool live = r < N;
    const float scale = rsqrtf((float)DH);
    float x[24];                                            // the residual stream of this (row, column group)
    if (live) load24_global(a.x + ((size_t)img * N + r) * D + 24 * g, x);
    else {
#pragma unroll
      for (int c = 0; c < 24; c++) x[c] = 0.f;
    }
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
    auto arrive = [&](int b) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[b]);
    };
    for (int bi = 0; bi < a.nblocks; bi++) {
      const NtVitBlock& P = a.blk[bi];
      const uint32_t ph = bi & 1;
      float v[24];
      // ---- u = LN(x) -> A planes                                       attention.py:22
      {
        float mean, rstd;
        row_stats(x, stat, r, g, mean, rstd);
#pragma unroll
        for (int c = 0; c < 24; c++)
          v[c] = live ? (x[c] - mean) * rstd * __ldg(P.ln_g + 24 * g + c) + __ldg(P.ln_b + 24 * g + c) : 0.f;
        store24_split(sA, sA + 2 * CH, r, 24 * g, v);
        arrive(FB_U);
      }
      // ---- q | k | v = acc + bias -> head-padded planes (+ fp32 copy for the backward / the layer API)
      for (int nt = 0; nt < 3; nt++) {
        mbar_wait(&bar[FB_QKVACC + nt], ph);
        tc_fence_after();
        ld24(lane_addr + NT * nt + 24 * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) v[c] = live ? v[c] + __ldg(P.bqkv + NT * nt + 24 * g + c) : 0.f;
        if (live) store24_global(P.qkv + ((size_t)img * N + r) * 3 * D + NT * nt + 24 * g, v);
        store24_split(sB + nt * 4 * CH, sB + nt * 4 * CH + 2 * CH, r, DHP * g, v);
        // the 8 padding columns of the head: P (heads 2, 3) and G were parked over these planes, so re-zero them every block
        const uint32_t zoff = plane_off(r, DHP * g + DH, CH);
        *reinterpret_cast<uint4*>(sB + nt * 4 * CH + zoff) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sB + nt * 4 * CH + 2 * CH + zoff) = make_uint4(0, 0, 0, 0);
      }
      arrive(FB_QKVRDY);
      // ---- P = softmax(S / sqrt(dh)) for head g                         attention.py:37-41
      {
        mbar_wait(&bar[FB_S + g], ph);
        if (g >= 2) mbar_wait(&bar[FB_S + 3], ph);          // heads 2, 3 park P in the q planes: every S product must be done
        tc_fence_after();
        uint32_t s0[32], s1[32];
        tmem_ld32(lane_addr + 64 * g, s0);
        tmem_ld32(lane_addr + 64 * g + 32, s1);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float t0 = j < N ? __uint_as_float(s0[j]) * scale : -INFINITY;
          const float t1 = 32 + j < N ? __uint_as_float(s1[j]) * scale : -INFINITY;
          s0[j] = __float_as_uint(t0);
          s1[j] = __float_as_uint(t1);
          mx = fmaxf(mx, fmaxf(t0, t1));
        }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float e0 = __expf(__uint_as_float(s0[j]) - mx), e1 = __expf(__uint_as_float(s1[j]) - mx);
          s0[j] = __float_as_uint(e0);
          s1[j] = __float_as_uint(e1);
          sum += e0 + e1;
        }
        const float inv = live ? __fdividef(1.0f, sum) : 0.f;
        uint8_t* ph_ = (g < 2 ? sA : sB) + (g & 1) * 2 * CH;
        float* pg = P.P ? P.P + (((size_t)img * H + g) * N + r) * N : nullptr;
#pragma unroll
        for (int j8 = 0; j8 < 8; j8++) {
          float w[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const int j = 8 * j8 + i;
            w[i] = live ? __uint_as_float(j < 32 ? s0[j & 31] : s1[j & 31]) * inv : 0.f;
          }
          store8_split(ph_, ph_ + CH, r, 8 * j8, CH, w);
          if (pg && live) {
#pragma unroll
            for (int i = 0; i < 8; i++)
              if (8 * j8 + i < N) pg[8 * j8 + i] = w[i];                 // the reference's atg__ (attention.py:41)
          }
        }
        arrive(FB_P + g);
      }
      // ---- h = x + O_g ; u1 = LN1(h)                                   attention.py:46-48
      {
        mbar_wait(&bar[FB_O + g], ph);
        tc_fence_after();
        ld24(lane_addr + 288 + DHP * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) x[c] = live ? x[c] + v[c] : 0.f;
        if (live) store24_global(P.h + ((size_t)img * N + r) * D + 24 * g, x);
        float mean, rstd;
        row_stats(x, stat, r, g, mean, rstd);
#pragma unroll
        for (int c = 0; c < 24; c++)
          v[c] = live ? (x[c] - mean) * rstd * __ldg(P.ln1_g + 24 * g + c) + __ldg(P.ln1_b + 24 * g + c) : 0.f;
        store24_split(sA, sA + 2 * CH, r, 24 * g, v);
        arrive(FB_U1);
      }
      // ---- pre = acc + b0 ; G = GELU(pre) -> planes (aliasing q | k)    attention.py:49-50
      for (int nt = 0; nt < 2; nt++) {
        mbar_wait(&bar[FB_F0 + nt], ph);
        tc_fence_after();
        ld24(lane_addr + NT * nt + 24 * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) v[c] = live ? v[c] + __ldg(P.b0 + NT * nt + 24 * g + c) : 0.f;
        if (live) store24_global(P.dgelu + ((size_t)img * N + r) * HD + NT * nt + 24 * g, v);   // pre-activation, kept for backward
#pragma unroll
        for (int c = 0; c < 24; c++) {
          float y, dy;
          gelu_both(v[c], y, dy);
          v[c] = live ? y : 0.f;
        }
        store24_split(sB, sB + 3 * CH, r, NT * nt + 24 * g, v);
      }
      arrive(FB_G);
      // ---- out = h + acc + b1                                           attention.py:51-53
      {
        mbar_wait(&bar[FB_F1], ph);
        tc_fence_after();
        ld24(lane_addr + 416 + 24 * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) x[c] = live ? x[c] + v[c] + __ldg(P.b1 + 24 * g + c) : 0.f;
        if (live) store24_global(P.out + ((size_t)img * N + r) * D + 24 * g, x);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc<512>(tbase);
  }
}

static bool supported(int N, int D_, int H_) { return N >= 1 && N <= ROWS && D_ == D && H_ == H; }

}  // namespace vtc
}  // namespace nt
using namespace nt;

extern "C" {

int nt_vit_tc_supported(int N, int D, int H, int* ok) {
  *ok = vtc::supported(N, D, H) ? 1 : 0;
  return 0;
}

int nt_vit_tc_wimg_bytes(int D, size_t* bytes) {
  (void)D;
  *bytes = (size_t)vtc::ALL_STAGES * vtc::STAGE;
  return 0;
}

static int vtc_fill(vtc::Args& a, const NtVitBlock* blocks, int nblocks, const char* who, int N, int D, int H) {
  NT_REQUIRE(vtc::supported(N, D, H), "%s: unsupported shape N=%d D=%d H=%d", who, N, D, H);
  NT_REQUIRE(nblocks >= 1 && nblocks <= vtc::MAXB, "%s: 1..%d blocks per call, got %d", who, vtc::MAXB, nblocks);
  for (int i = 0; i < nblocks; i++) a.blk[i] = blocks[i];
  a.nblocks = nblocks;
  a.N = N;
  return 0;
}

int nt_vit_tc_prepare(const NtVitBlock* blocks, int nblocks, int D, int H) {
  NT_ENTRY();
  vtc::Args a{};
  if (int rc = vtc_fill(a, blocks, nblocks, "nt_vit_tc_prepare", 1, D, H)) return rc;
  NT_CHECK(launch_k(vtc::vit_tc_wprep_kernel, dim3(ceil_div(vtc::ALL_STAGES * vtc::NT * 8, 256), nblocks), dim3(256), 0, a));
  NT_LAUNCH_OK();
  return 0;
}

int nt_vit_tc_fwd(const NtVitBlock* blocks, int nblocks, const float* x, int B, int N, int D, int H, int prepared) {
  NT_ENTRY();
  vtc::Args a{};
  if (int rc = vtc_fill(a, blocks, nblocks, "nt_vit_tc_fwd", N, D, H)) return rc;
  if (B == 0) return 0;
  a.x = x;
  if (!prepared) {
    NT_CHECK(launch_k(vtc::vit_tc_wprep_kernel, dim3(ceil_div(vtc::ALL_STAGES * vtc::NT * 8, 256), nblocks), dim3(256), 0, a));
    NT_LAUNCH_OK();
  }
  static bool attr = false;
  if (!attr) {
    NT_CHECK(cudaFuncSetAttribute(vtc::vit_tc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtc::SMEM_BYTES));
    attr = true;
  }
  NT_CHECK(launch_k(vtc::vit_tc_fwd_kernel, dim3(B), dim3(vtc::THREADS), vtc::SMEM_BYTES, a));
  NT_LAUNCH_OK();
  return 0;
}

}  // extern "C"
