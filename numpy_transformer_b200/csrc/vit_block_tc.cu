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
#include <cstdlib>

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
  long long* dbg;       // optional: phase time stamps (clock64) of CTA 0, [role][block][event]
  int dbg_flags;        // experiments (results become WRONG): 1 = token warps do not wait for the weight-gradient flush
};
__device__ __forceinline__ void stamp(const Args& a, int role, int bi, int ev) {
  if (a.dbg && blockIdx.x == 0) a.dbg[(role * MAXB + bi) * 16 + ev] = clock64();
}

constexpr int D = 96, H = 4, DH = 24, DHP = 32, HD = 192;
constexpr int ROWS = 64;                       // token rows of a plane chunk
constexpr uint32_t CH = ROWS * 128;            // bytes of one 64-column chunk of a plane (8 KB)
constexpr int NT = 96;                         // rows (output features) of a weight tile
constexpr uint32_t WPLANE = NT * 128;          // 12 KB
constexpr uint32_t STAGE = 2 * WPLANE;         // hi + lo
constexpr int FWD_STAGES = 13, BWD_STAGES = 13, ALL_STAGES = FWD_STAGES + BWD_STAGES;
constexpr int NSTAGE = 3;
constexpr int THREADS = 448;                   // 14 warps: epilogue {0,1,4,5,8,9,12,13}, loader 2, MMA 3
constexpr int EPI_THREADS = 256;

// shared memory map (bytes)
constexpr uint32_t OFF_A = 0;                        // 4 chunks: U / U1 (hi 2, lo 2)  |  P0 hi, P0 lo, P1 hi, P1 lo
constexpr uint32_t OFF_B = OFF_A + 4 * CH;           // 12 chunks: Q hi 2, Q lo 2, K hi 2, K lo 2, V hi 2, V lo 2  |  G hi 3, G lo 3
constexpr uint32_t OFF_RING = OFF_B + 12 * CH;
constexpr uint32_t OFF_STAT = OFF_RING + NSTAGE * STAGE;      // [2][64][4] floats
// the block's small parameter vectors, copied once per block by the loader (double-buffered):
// ln_g 0 | ln_b 96 | bqkv 192 | ln1_g 480 | ln1_b 576 | b0 672 | b1 864   (960 floats)
constexpr int PAR_FLOATS = 960, PV_LNG = 0, PV_LNB = 96, PV_BQKV = 192, PV_LN1G = 480, PV_LN1B = 576, PV_B0 = 672, PV_B1 = 864;
constexpr uint32_t OFF_PAR = OFF_STAT + 2 * 64 * 4 * 4;
constexpr uint32_t OFF_BAR = OFF_PAR + 2 * PAR_FLOATS * 4;
constexpr int NBAR = 2 * NSTAGE + 26;
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
// "Tiled" fp32 saves: a per-image [N][W] tensor is kept as [W/4][N][4] (16-byte unit u of every row r contiguous over r),
// so that the 32 lanes of a warp — 32 consecutive rows — store / load 512 contiguous bytes per instruction instead of 32
// separate sectors.  Only this kernel pair reads them; the host side un-tiles on demand (fused_vit.untile).
__device__ __forceinline__ void store24_tiled(float* img_base, int N, int r, int u0, const float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 6; j++)
    *reinterpret_cast<float4*>(img_base + ((size_t)(u0 + j) * N + r) * 4) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
}
__device__ __forceinline__ void load24_tiled(const float* img_base, int N, int r, int u0, float (&v)[24]) {
#pragma unroll
  for (int j = 0; j < 6; j++) {
    const float4 t = *reinterpret_cast<const float4*>(img_base + ((size_t)(u0 + j) * N + r) * 4);
    v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
  }
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
       FB_G = 20, FB_F1 = 21, FB_PAR = 22 /*2*/, FB_COUNT = 24 };

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
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 26);
  float* par = reinterpret_cast<float*>(smem + OFF_PAR);

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
    mbar_init(&bar[FB_PAR], 1); mbar_init(&bar[FB_PAR + 1], 1);
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
          if (i == 0) {
            // this block's bias / LayerNorm vectors -> shared memory.  The buffer was last used two blocks ago: the ring slot
            // just obtained proves the MMAs are within NSTAGE tiles of here, i.e. the token warps are past that block
            const NtVitBlock& P = a.blk[bi];
            float* pb = par + (bi & 1) * PAR_FLOATS;
            uint64_t* pbar = &bar[FB_PAR + (bi & 1)];
            mbar_expect_tx(pbar, PAR_FLOATS * 4);
            bulk_g2s(pb + PV_LNG, P.ln_g, D * 4, pbar);    bulk_g2s(pb + PV_LNB, P.ln_b, D * 4, pbar);
            bulk_g2s(pb + PV_BQKV, P.bqkv, 3 * D * 4, pbar);
            bulk_g2s(pb + PV_LN1G, P.ln1_g, D * 4, pbar);  bulk_g2s(pb + PV_LN1B, P.ln1_b, D * 4, pbar);
            bulk_g2s(pb + PV_B0, P.b0, HD * 4, pbar);      bulk_g2s(pb + PV_B1, P.b1, D * 4, pbar);
          }
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
        stamp(a, 1, bi, 0);
        mbar_wait(&bar[FB_U], ph);
        tc_fence_after();
        stamp(a, 1, bi, 1);
        for (int nt = 0; nt < 3; nt++) {
          for (int kc = 0; kc < 2; kc++) {
            const uint32_t st = stage_wait();
            mma_stage(aA + kc * CH, aA + 2 * CH + kc * CH, st, kc ? 2 : 4, tbase + NT * nt, id_lin, kc == 0);
            stage_done();
          }
          mma_commit(&bar[FB_QKVACC + nt]);
        }
        // ---- S_h = Q_h K_h^T                                           attention.py:35-36
        stamp(a, 1, bi, 2);
        mbar_wait(&bar[FB_QKVRDY], ph);
        tc_fence_after();
        stamp(a, 1, bi, 3);
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
        stamp(a, 1, bi, 4);
        mbar_wait(&bar[FB_U1], ph);
        tc_fence_after();
        stamp(a, 1, bi, 5);
        for (int nt = 0; nt < 2; nt++) {
          for (int kc = 0; kc < 2; kc++) {
            const uint32_t st = stage_wait();
            mma_stage(aA + kc * CH, aA + 2 * CH + kc * CH, st, kc ? 2 : 4, tbase + NT * nt, id_lin, kc == 0);
            stage_done();
          }
          mma_commit(&bar[FB_F0 + nt]);
        }
        // ---- y = GELU(pre) W1                                          attention.py:51
        stamp(a, 1, bi, 6);
        mbar_wait(&bar[FB_G], ph);
        tc_fence_after();
        stamp(a, 1, bi, 7);
        for (int kc = 0; kc < 3; kc++) {
          const uint32_t st = stage_wait();
          mma_stage(aB + kc * CH, aB + 3 * CH + kc * CH, st, 4, tbase + 416, id_lin, kc == 0);
          stage_done();
        }
        mma_commit(&bar[FB_F1]);
        stamp(a, 1, bi, 8);
      }
    }
    __syncwarp();
  } else if (is_epi) {
    // ===================== epilogue warps: one token row per thread, 24 columns (= one head) per column group ========
    // The phases are ROLLED loops over 8-column units: a fully unrolled version is ~180 KB of straight-line code and runs
    // at the speed of the instruction cache misses (ncu: ICC hit rate 81 %, profiles/r02_ncu_full_vit_tc_bwd_v1_summary.csv).
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
    // y[c] = (x[c] - mean) * rstd * gamma[c] + beta[c] -> split planes at columns 24 g + c (unrolled: x lives in registers)
    auto layernorm_to_planes = [&](const float* gam, const float* bet) {
      float mean, rstd;
      row_stats(x, stat, r, g, mean, rstd);
#pragma unroll
      for (int j = 0; j < 3; j++) {
        float w[8];
        const float4 g0 = *reinterpret_cast<const float4*>(gam + 24 * g + 8 * j), g1 = *reinterpret_cast<const float4*>(gam + 24 * g + 8 * j + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(bet + 24 * g + 8 * j), b1 = *reinterpret_cast<const float4*>(bet + 24 * g + 8 * j + 4);
        const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = live ? (x[8 * j + i] - mean) * rstd * gg[i] + bb[i] : 0.f;
        store8_split(sA, sA + 2 * CH, r, 24 * g + 8 * j, CH, w);
      }
    };
    // eight accumulator columns starting at TMEM column `col` (+ vector `add8` from shared memory)
    auto ld8_add = [&](uint32_t col, const float* add8, float (&w)[8]) {
      uint32_t t[8];
      tmem_ld8(lane_addr + col, t);
      const float4 b0 = *reinterpret_cast<const float4*>(add8), b1 = *reinterpret_cast<const float4*>(add8 + 4);
      tmem_wait_ld();
      w[0] = __uint_as_float(t[0]) + b0.x; w[1] = __uint_as_float(t[1]) + b0.y; w[2] = __uint_as_float(t[2]) + b0.z; w[3] = __uint_as_float(t[3]) + b0.w;
      w[4] = __uint_as_float(t[4]) + b1.x; w[5] = __uint_as_float(t[5]) + b1.y; w[6] = __uint_as_float(t[6]) + b1.z; w[7] = __uint_as_float(t[7]) + b1.w;
      if (!live) {
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = 0.f;
      }
    };
    auto store8_tiled = [&](float* img_base, int u, const float (&w)[8]) {     // units u, u + 1 of row r
      *reinterpret_cast<float4*>(img_base + ((size_t)u * N + r) * 4) = make_float4(w[0], w[1], w[2], w[3]);
      *reinterpret_cast<float4*>(img_base + ((size_t)(u + 1) * N + r) * 4) = make_float4(w[4], w[5], w[6], w[7]);
    };
    for (int bi = 0; bi < a.nblocks; bi++) {
      const NtVitBlock& P = a.blk[bi];
      const uint32_t ph = bi & 1;
      const float* pv = par + (bi & 1) * PAR_FLOATS;
      const bool t0 = tid == 0;
      if (t0) stamp(a, 0, bi, 0);
      mbar_wait(&bar[FB_PAR + (bi & 1)], (bi >> 1) & 1);
      // ---- u = LN(x) -> A planes                                       attention.py:22
      layernorm_to_planes(pv + PV_LNG, pv + PV_LNB);
      arrive(FB_U);
      if (t0) stamp(a, 0, bi, 1);
      // ---- q | k | v = acc + bias -> head-padded planes (+ fp32 copy, tiled, for the backward / the layer API)
      {
        float* qkv_img = P.qkv + (size_t)img * N * 3 * D;
#pragma unroll 1
        for (int it = 0; it < 9; it++) {
          const int nt = it / 3, j = it - 3 * nt;
          if (j == 0) {
            mbar_wait(&bar[FB_QKVACC + nt], ph);
            tc_fence_after();
            if (t0) stamp(a, 0, bi, 2 + nt);
          }
          float w[8];
          ld8_add(NT * nt + 24 * g + 8 * j, pv + PV_BQKV + NT * nt + 24 * g + 8 * j, w);
          if (live) store8_tiled(qkv_img, 24 * nt + 6 * g + 2 * j, w);
          uint8_t* hi = sB + nt * 4 * CH;
          store8_split(hi, hi + 2 * CH, r, DHP * g + 8 * j, CH, w);
          if (j == 2) {     // the 8 padding columns of the head: P (heads 2, 3) and G were parked over these planes
            const uint32_t zoff = plane_off(r, DHP * g + DH, CH);
            *reinterpret_cast<uint4*>(hi + zoff) = make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(hi + 2 * CH + zoff) = make_uint4(0, 0, 0, 0);
          }
        }
      }
      arrive(FB_QKVRDY);
      if (t0) stamp(a, 0, bi, 5);
      // ---- P = softmax(S / sqrt(dh)) for head g: three passes over the 64 key columns, 16 at a time  attention.py:37-41
      {
        mbar_wait(&bar[FB_S + g], ph);
        if (g >= 2) mbar_wait(&bar[FB_S + 3], ph);          // heads 2, 3 park P in the q planes: every S product must be done
        tc_fence_after();
        if (t0) stamp(a, 0, bi, 6);
        // the whole row of scores in registers: one TMEM round trip (reading it in pieces costs a load latency per piece)
        uint32_t s0[32], s1[32];
        tmem_ld32(lane_addr + 64 * g, s0);
        tmem_ld32(lane_addr + 64 * g + 32, s1);
        tmem_wait_ld();
        float mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float t0_ = j < N ? __uint_as_float(s0[j]) * scale : -INFINITY;
          const float t1_ = 32 + j < N ? __uint_as_float(s1[j]) * scale : -INFINITY;
          s0[j] = __float_as_uint(t0_);
          s1[j] = __float_as_uint(t1_);
          mx = fmaxf(mx, fmaxf(t0_, t1_));
        }
        float sum = 0.f, sum1 = 0.f;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          const float e0 = __expf(__uint_as_float(s0[j]) - mx), e1 = __expf(__uint_as_float(s1[j]) - mx);
          s0[j] = __float_as_uint(e0);
          s1[j] = __float_as_uint(e1);
          sum += e0;
          sum1 += e1;
        }
        const float inv = live ? __fdividef(1.0f, sum + sum1) : 0.f;
        uint8_t* ph_ = (g < 2 ? sA : sB) + (g & 1) * 2 * CH;
        float* pg = P.P ? P.P + ((size_t)img * H + g) * (size_t)((N + 3) / 4) * N * 4 : nullptr;   // tiled [ceil(N/4)][N][4]
#pragma unroll
        for (int j8 = 0; j8 < 8; j8++) {
          float w[8];
#pragma unroll
          for (int i = 0; i < 8; i++) w[i] = __uint_as_float(8 * j8 + i < 32 ? s0[(8 * j8 + i) & 31] : s1[(8 * j8 + i) & 31]) * inv;
          store8_split(ph_, ph_ + CH, r, 8 * j8, CH, w);
          if (pg && live) {                                             // the reference's atg__ (attention.py:41)
            if (8 * j8 < N) *reinterpret_cast<float4*>(pg + ((size_t)(2 * j8) * N + r) * 4) = make_float4(w[0], w[1], w[2], w[3]);
            if (8 * j8 + 4 < N) *reinterpret_cast<float4*>(pg + ((size_t)(2 * j8 + 1) * N + r) * 4) = make_float4(w[4], w[5], w[6], w[7]);
          }
        }
        arrive(FB_P + g);
        if (t0) stamp(a, 0, bi, 7);
      }
      // ---- h = x + O_g ; u1 = LN1(h)                                   attention.py:46-48
      {
        mbar_wait(&bar[FB_O + g], ph);
        tc_fence_after();
        if (t0) stamp(a, 0, bi, 8);
        float v[24];
        ld24(lane_addr + 288 + DHP * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) x[c] = live ? x[c] + v[c] : 0.f;
        if (live) store24_tiled(P.h + (size_t)img * N * D, N, r, 6 * g, x);
        layernorm_to_planes(pv + PV_LN1G, pv + PV_LN1B);
        arrive(FB_U1);
        if (t0) stamp(a, 0, bi, 9);
      }
      // ---- pre = acc + b0 ; G = GELU(pre) -> planes (aliasing q | k)    attention.py:49-50
      {
        float* pre_img = P.dgelu + (size_t)img * N * HD;
#pragma unroll 1
        for (int it = 0; it < 6; it++) {
          const int nt = it / 3, j = it - 3 * nt;
          if (j == 0) {
            mbar_wait(&bar[FB_F0 + nt], ph);
            tc_fence_after();
            if (t0) stamp(a, 0, bi, 10 + nt);
          }
          float w[8];
          ld8_add(NT * nt + 24 * g + 8 * j, pv + PV_B0 + NT * nt + 24 * g + 8 * j, w);
          if (live) store8_tiled(pre_img, 24 * nt + 6 * g + 2 * j, w);      // pre-activation, kept for the backward
#pragma unroll
          for (int i = 0; i < 8; i++) {
            float y, dy;
            gelu_both(w[i], y, dy);
            w[i] = live ? y : 0.f;
          }
          store8_split(sB, sB + 3 * CH, r, NT * nt + 24 * g + 8 * j, CH, w);
        }
      }
      arrive(FB_G);
      if (t0) stamp(a, 0, bi, 12);
      // ---- out = h + acc + b1                                           attention.py:51-53
      {
        mbar_wait(&bar[FB_F1], ph);
        tc_fence_after();
        if (t0) stamp(a, 0, bi, 13);
        float v[24];
        ld24(lane_addr + 416 + 24 * g, v);
#pragma unroll
        for (int c = 0; c < 24; c++) x[c] = live ? x[c] + v[c] + pv[PV_B1 + 24 * g + c] : 0.f;
        if (live) {
          // the stack's output in the caller's row-major layout; block outputs in between are only read by the backward
          if (bi == a.nblocks - 1) store24_global(P.out + ((size_t)img * N + r) * D + 24 * g, x);
          else store24_tiled(P.out + (size_t)img * N * D, N, r, 6 * g, x);
        }
        if (t0) stamp(a, 0, bi, 14);
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

// ================================================================================================ backward
// Warp roles (16 warps): token warps {0,1,4,5,8,9,12,13} (one token row per thread, column group g = head g), loader 2,
// critical-path MMA issuer 3, weight-gradient ("flush") warps {6,7,10,11} on TMEM lanes 64..127, flush MMA issuer 14.
//
// Weight gradients dW = X^T dY are contracted over the image's tokens with the activation planes as MN-major operands
// (no transposes).  Their M index is the input FEATURE; the A descriptor starts one 64-feature chunk early so that the
// wanted 64 features land in accumulator rows 64..127 = the TMEM lanes of warps that are NOT on the critical path, which
// add them to the gradient arena with red.global.add.v4 (the reference's `+=` into *_delta, fullconnect.py:118-122).  A
// constant 1 in the first padding column of U / U1 makes row 96 of dWqkv / dW0 the bias gradient.
constexpr int BTHREADS = 512;
constexpr int BNSTAGE = 2;
constexpr uint32_t BOFF_D = 0;                              // 4 chunks
constexpr uint32_t BOFF_X = 4 * CH;                         // 16 chunks
constexpr uint32_t BOFF_RING = 20 * CH;
constexpr uint32_t BOFF_STAT = BOFF_RING + BNSTAGE * STAGE; // [2][64][4] floats
constexpr uint32_t BOFF_ACC = BOFF_STAT + 2 * 64 * 4 * 4;       // [5 quantities][96] floats: column sums (bias / LayerNorm gradients)
constexpr uint32_t BOFF_PAR = BOFF_ACC + 5 * 96 * 4;        // [2 buffers][ln_g | ln_b | ln1_g | ln1_b] floats, copied by the loader
constexpr int BPAR_FLOATS = 4 * 96;
constexpr uint32_t STG_PITCH = 80;                              // bytes per staged row: 16 floats + 16 (conflict-free 16-byte stores)
constexpr uint32_t BOFF_STG = BOFF_PAR + 2 * BPAR_FLOATS * 4;   // [4 flush warps][32 rows][STG_PITCH]: transposition tiles of the flush
constexpr uint32_t BOFF_BAR = BOFF_STG + 4 * 32 * STG_PITCH;
constexpr int BNBAR = 2 * BNSTAGE + 49;
constexpr uint32_t BSMEM_BYTES = BOFF_BAR + BNBAR * 8 + 16;
static_assert(BSMEM_BYTES <= 232448, "shared memory budget (backward)");
// chunk indices inside the operand area (chunk c lives at smem + c * CH)
constexpr int C_DH = 0, C_DL = 2;                           // d / dO: hi chunks 0,1  lo chunks 2,3
constexpr int C_GH = 4, C_GL = 7, C_PH = 10, C_PL = 13, C_U1H = 16, C_U1L = 18;     // MLP phase: G, dpre, LN1(h)
constexpr int C_QH = 4, C_QL = 10;                          // attention phase: [q c0 c1 | k c0 c1 | v c0 c1] hi, then lo
// P / dS of head h: hi chunk ps_chunk(h), lo the next one.  Heads 2, 3 borrow the weight ring, which is idle between the
// dU1 and the dU products (the loader holds the dU tiles back until the attention phase is over).
__host__ __device__ constexpr int ps_chunk(int h) { return h < 2 ? 16 + 2 * h : 20 + 2 * (h - 2); }
constexpr int C_UH = 16, C_UL = 18;                         // LN(x) for dWqkv (after the attention phase)
enum { BB_D = 0, BB_DGACC = 1 /*2*/, BB_GDP = 3, BB_DU1ACC = 4, BB_ATT = 5, BB_SACC = 6 /*4*/, BB_SCONS = 10 /*4*/, BB_DPACC = 14 /*4*/,
       BB_P = 18 /*4*/, BB_DVACC = 22 /*4*/, BB_DS = 26 /*4*/, BB_DQKACC = 30 /*4*/, BB_DQKV = 34, BB_U = 35, BB_DUACC = 36,
       BB_DW1_DONE = 37, BB_DW0_DONE = 38, BB_DWQKV_DONE = 39, BB_FFULL = 40 /*2*/, BB_FFREE = 42 /*2*/, BB_ACCRDY = 44,
       BB_ACCFREE = 45 /*2*/, BB_PAR = 47 /*2*/, BB_COUNT = 49 };
constexpr int PIECES = 24;                                  // weight-gradient pieces per block: dW1 6, dW0 6, dWqkv 12

// column sums over the 32 rows of a warp of 24 per-thread values: reduce-scatter (the lanes trade halves of their columns),
// then two full butterfly steps.  Afterwards every lane holds the sums of columns c0, c0+1, c0+2 with
// c0 = 12*bit4(lane) + 6*bit3(lane) + 3*bit2(lane); the lanes with (lane & 3) == 0 cover all 24 columns once.
__device__ __forceinline__ void colsum24(const float (&v)[24], int lane, float (&out)[3], int& c0) {
  float a[12], b[6];
  const bool u16 = lane & 16, u8 = lane & 8, u4 = lane & 4;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const float send = u16 ? v[i] : v[12 + i], keep = u16 ? v[12 + i] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 6; i++) {
    const float send = u8 ? a[i] : a[6 + i], keep = u8 ? a[6 + i] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const float send = u4 ? b[i] : b[3 + i], keep = u4 ? b[3 + i] : b[i];
    float t = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    out[i] = t;
  }
  c0 = (u16 ? 12 : 0) + (u8 ? 6 : 0) + (u4 ? 3 : 0);
}
// dst[c] += sum over the warp's rows of v[c]; dst is a SHARED-memory accumulator (the global reds are issued later by a
// warp that is not on the critical path: 100 CTAs adding to the same 96 addresses made the token warps wait)
__device__ __forceinline__ void colsum24_atomic(const float (&v)[24], int lane, float* dst) {
  float o[3];
  int c0;
  colsum24(v, lane, o, c0);
  if ((lane & 3) == 0) {
    atomicAdd(dst + c0, o[0]);
    atomicAdd(dst + c0 + 1, o[1]);
    atomicAdd(dst + c0 + 2, o[2]);
  }
}
__device__ __forceinline__ void red4(float* p, float x, float y, float z, float w) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
// row-wise sums of two per-thread partials over the 4 column groups (LayerNorm backward)
__device__ __forceinline__ void row_sum2(float p1, float p2, float* stat, int r, int g, float& s1, float& s2) {
  stat[r * 4 + g] = p1;
  stat[256 + r * 4 + g] = p2;
  named_bar_sync(1, EPI_THREADS);
  const float4 t = *reinterpret_cast<const float4*>(stat + r * 4), u = *reinterpret_cast<const float4*>(stat + 256 + r * 4);
  s1 = t.x + t.y + t.z + t.w;
  s2 = u.x + u.y + u.z + u.w;
  named_bar_sync(1, EPI_THREADS);            // the buffer may be rewritten right away
}
// a 16-byte unit [1, 0, .., 0] (bf16) at column c0 of row r: the "ones" feature that turns a weight-gradient row into the bias gradient
__device__ __forceinline__ void store_ones_unit(uint8_t* hi, uint8_t* lo, int r, int c0, bool one) {
  const uint32_t off = plane_off(r, c0, CH);
  *reinterpret_cast<uint4*>(hi + off) = make_uint4(one ? 0x00003F80u : 0u, 0, 0, 0);
  *reinterpret_cast<uint4*>(lo + off) = make_uint4(0, 0, 0, 0);
}

__global__ void __launch_bounds__(BTHREADS, 1) vit_tc_bwd_kernel(Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem_raw) & 1023u) __trap();
  float* stat = reinterpret_cast<float*>(smem + BOFF_STAT);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + BOFF_BAR);
  uint64_t* empty = full + BNSTAGE;
  uint64_t* bar = empty + BNSTAGE;
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + 49);
  float* par = reinterpret_cast<float*>(smem + BOFF_PAR);
  auto chunk = [&](int c) -> uint8_t* { return smem + (uint32_t)c * CH; };

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, img = blockIdx.x, L = a.nblocks;
  const bool is_tok = (warp & 3) < 2;
  const int q = warp & 3, g = warp >> 2, r = 32 * q + lane;

  if (tid == 0) {
    for (int s = 0; s < BNSTAGE; s++) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int i = 0; i < BB_COUNT; i++) mbar_init(&bar[i], 1);
    const int eights[] = {BB_D, BB_GDP, BB_ATT, BB_DQKV, BB_U, BB_ACCRDY};
    for (int i = 0; i < 6; i++) mbar_init(&bar[eights[i]], 8);
    for (int i = 0; i < 4; i++) { mbar_init(&bar[BB_SCONS + i], 2); mbar_init(&bar[BB_P + i], 2); mbar_init(&bar[BB_DS + i], 2); }
    mbar_init(&bar[BB_FFREE], 4); mbar_init(&bar[BB_FFREE + 1], 4);
    mbar_fence_init();
  }
  for (uint32_t i = tid * 16; i < 20 * CH; i += BTHREADS * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  float* accs = reinterpret_cast<float*>(smem + BOFF_ACC);
  for (int i = tid; i < 5 * 96; i += BTHREADS) accs[i] = 0.f;
  if (warp == 3) tmem_alloc<512>(tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  const uint32_t sm0 = smem_u32(smem);
  auto caddr = [&](int c) -> uint32_t { return sm0 + (uint32_t)c * CH; };
  pdl_wait();

  if (warp == 2) {
    // ===================== loader =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int bi = L - 1; bi >= 0; bi--) {
        const uint8_t* wimg = reinterpret_cast<const uint8_t*>(a.blk[bi].wfrag) + (size_t)FWD_STAGES * STAGE;
        for (int i = 0; i < BWD_STAGES; i++, it++) {
          const int s = it % BNSTAGE;
          if (it >= BNSTAGE) mbar_wait(&empty[s], ((it / BNSTAGE) - 1) & 1);
          if (i == 7) mbar_wait(&bar[BB_DQKV], (L - 1 - bi) & 1);     // heads 2, 3 keep P / dS in the ring until here
          if (i == 0) {      // the block's LayerNorm vectors -> shared memory (buffer last used two iterations ago, see the forward)
            const NtVitBlock& P = a.blk[bi];
            const int nb = (L - 1 - bi) & 1;
            float* pb = par + nb * BPAR_FLOATS;
            mbar_expect_tx(&bar[BB_PAR + nb], BPAR_FLOATS * 4);
            bulk_g2s(pb, P.ln_g, D * 4, &bar[BB_PAR + nb]);            bulk_g2s(pb + 96, P.ln_b, D * 4, &bar[BB_PAR + nb]);
            bulk_g2s(pb + 192, P.ln1_g, D * 4, &bar[BB_PAR + nb]);     bulk_g2s(pb + 288, P.ln1_b, D * 4, &bar[BB_PAR + nb]);
          }
          mbar_expect_tx(&full[s], STAGE);
          bulk_g2s(smem + BOFF_RING + s * STAGE, wimg + (size_t)i * STAGE, STAGE, &full[s]);
        }
      }
    }
    __syncwarp();
  } else if (warp == 3) {
    // ===================== critical-path MMA issuer =====================
    if (lane == 0) {
      const uint32_t id_lin = idesc_bf16(128, NT, 0, 0), id_s = idesc_bf16(128, 64, 0, 0);
      const uint32_t id_dv = idesc_bf16(128, DHP, 1, 1), id_dq = idesc_bf16(128, DHP, 0, 1);
      uint32_t it = 0;
      auto stage_wait = [&]() -> uint32_t {
        const int s = it % BNSTAGE;
        mbar_wait(&full[s], (it / BNSTAGE) & 1);
        tc_fence_after();
        return sm0 + BOFF_RING + s * STAGE;
      };
      auto stage_done = [&]() { mma_commit(&empty[it % BNSTAGE]); it++; };
      auto wait = [&](int b, uint32_t ph) { mbar_wait(&bar[b], ph); tc_fence_after(); };
      for (int n = 0; n < L; n++) {
        const uint32_t ph = n & 1;
        // ---- dG = d W1^T                                               attention.py:58
        wait(BB_D, ph);
        for (int nt = 0; nt < 2; nt++) {
          for (int kc = 0; kc < 2; kc++) {
            const uint32_t st = stage_wait();
            mma_stage(caddr(C_DH + kc), caddr(C_DL + kc), st, kc ? 2 : 4, tbase + NT * nt, id_lin, kc == 0);
            stage_done();
          }
          mma_commit(&bar[BB_DGACC + nt]);
        }
        // ---- dU1 = dpre W0^T                                           attention.py:60
        wait(BB_GDP, ph);
        for (int kc = 0; kc < 3; kc++) {
          const uint32_t st = stage_wait();
          mma_stage(caddr(C_PH + kc), caddr(C_PL + kc), st, 4, tbase + 192, id_lin, kc == 0);
          stage_done();
        }
        mma_commit(&bar[BB_DU1ACC]);
        // ---- attention backward                                        attention.py:63-84
        wait(BB_ATT, ph);
        auto hoff = [&](int h) -> uint32_t { return (uint32_t)(h >> 1) * CH + (uint32_t)(h & 1) * 64u; };
        const uint32_t Qh = caddr(C_QH), Ql = caddr(C_QL), Kh = caddr(C_QH + 2), Kl = caddr(C_QL + 2), Vh = caddr(C_QH + 4), Vl = caddr(C_QL + 4);
        const uint32_t Oh = caddr(C_DH), Ol = caddr(C_DL);
        for (int h = 0; h < H; h++) {                         // S_h = Q_h K_h^T (recomputed)
          const uint32_t off = hoff(h);
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {
            const uint32_t o = off + ks * KMAJ_STEP;
            mma_bf16(tbase + 64 * h, desc_kmajor(Ql + o), desc_kmajor(Kh + o), id_s, ks ? 1u : 0u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Qh + o), desc_kmajor(Kl + o), id_s, 1u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Qh + o), desc_kmajor(Kh + o), id_s, 1u);
          }
          mma_commit(&bar[BB_SACC + h]);
        }
        auto dP = [&](int h) {
          const uint32_t off = hoff(h);
          wait(BB_SCONS + h, ph);                             // the softmax threads hold S_h in registers: its columns are free
#pragma unroll
          for (int ks = 0; ks < 2; ks++) {                    // dP_h = dO_h V_h^T  (attention.py:74)
            const uint32_t o = off + ks * KMAJ_STEP;
            mma_bf16(tbase + 64 * h, desc_kmajor(Ol + o), desc_kmajor(Vh + o), id_s, ks ? 1u : 0u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Oh + o), desc_kmajor(Vl + o), id_s, 1u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Oh + o), desc_kmajor(Vh + o), id_s, 1u);
          }
          mma_commit(&bar[BB_DPACC + h]);
        };
        auto dV = [&](int h) {
          const uint32_t off = hoff(h);
          wait(BB_P + h, ph);
          const uint32_t Ph = caddr(ps_chunk(h)), Pl = Ph + CH;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {                    // dV_h = P_h^T dO_h  (attention.py:76): both operands MN-major
            const uint32_t pa = ks * MNMAJ_STEP, ob = off + ks * MNMAJ_STEP;
            mma_bf16(tbase + 256 + DHP * h, desc_mnmajor(Pl + pa, CH), desc_mnmajor(Oh + ob, CH), id_dv, ks ? 1u : 0u);
            mma_bf16(tbase + 256 + DHP * h, desc_mnmajor(Ph + pa, CH), desc_mnmajor(Ol + ob, CH), id_dv, 1u);
            mma_bf16(tbase + 256 + DHP * h, desc_mnmajor(Ph + pa, CH), desc_mnmajor(Oh + ob, CH), id_dv, 1u);
          }
          mma_commit(&bar[BB_DVACC + h]);
        };
        auto dQ_dK = [&](int h) {
          const uint32_t off = hoff(h);
          wait(BB_DS + h, ph);
          const uint32_t Sh = caddr(ps_chunk(h)), Sl = Sh + CH;
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {                    // dQ_h = dS_h K_h  (attention.py:78), dS already carries 1/sqrt(dh)
            const uint32_t kb = off + ks * MNMAJ_STEP;
            mma_bf16(tbase + 64 * h, desc_kmajor(Sl + ks * KMAJ_STEP), desc_mnmajor(Kh + kb, CH), id_dq, ks ? 1u : 0u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Sh + ks * KMAJ_STEP), desc_mnmajor(Kl + kb, CH), id_dq, 1u);
            mma_bf16(tbase + 64 * h, desc_kmajor(Sh + ks * KMAJ_STEP), desc_mnmajor(Kh + kb, CH), id_dq, 1u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ks++) {                    // dK_h = dS_h^T Q_h  (attention.py:79)
            const uint32_t sa = ks * MNMAJ_STEP, qb = off + ks * MNMAJ_STEP;
            mma_bf16(tbase + 64 * h + 32, desc_mnmajor(Sl + sa, CH), desc_mnmajor(Qh + qb, CH), id_dv, ks ? 1u : 0u);
            mma_bf16(tbase + 64 * h + 32, desc_mnmajor(Sh + sa, CH), desc_mnmajor(Ql + qb, CH), id_dv, 1u);
            mma_bf16(tbase + 64 * h + 32, desc_mnmajor(Sh + sa, CH), desc_mnmajor(Qh + qb, CH), id_dv, 1u);
          }
          mma_commit(&bar[BB_DQKACC + h]);
        };
        for (int h = 0; h < H; h++) dP(h);
        for (int h = 0; h < H; h++) dV(h);
        for (int h = 0; h < H; h++) dQ_dK(h);
        // ---- dU = dqkv Wqkv^T                                          attention.py:85
        wait(BB_DQKV, ph);
        for (int kc = 0; kc < 6; kc++) {
          const uint32_t st = stage_wait();
          mma_stage(caddr(C_QH + kc), caddr(C_QL + kc), st, 4, tbase, id_lin, kc == 0);
          stage_done();
        }
        mma_commit(&bar[BB_DUACC]);
      }
    }
    __syncwarp();
  } else if (warp == 14) {
    // ===================== weight-gradient MMA issuer =====================
    if (lane == 0) {
      uint32_t piece = 0;
      auto wait = [&](int b, uint32_t ph) { mbar_wait(&bar[b], ph); tc_fence_after(); };
      // one piece: accumulator rows 64..127 <- features [64 f, 64 f + 64) of the A plane, N columns of the B plane chunk
      auto issue = [&](uint32_t a_hi, uint32_t a_lo, int f, uint32_t b_hi, uint32_t b_lo, int ncols) {
        const int buf = piece & 1;
        if (piece >= 2) wait(BB_FFREE + buf, ((piece >> 1) - 1) & 1);
        const uint32_t idesc = idesc_bf16(128, ncols, 1, 1), tacc = tbase + 384 + 64 * buf;
        const uint32_t ah = a_hi + (uint32_t)(f - 1) * CH, al = a_lo + (uint32_t)(f - 1) * CH;
#pragma unroll
        for (int ks = 0; ks < 4; ks++) {
          const uint32_t o = ks * MNMAJ_STEP;
          mma_bf16(tacc, desc_mnmajor(al + o, CH), desc_mnmajor(b_hi + o, CH), idesc, ks ? 1u : 0u);
          mma_bf16(tacc, desc_mnmajor(ah + o, CH), desc_mnmajor(b_lo + o, CH), idesc, 1u);
          mma_bf16(tacc, desc_mnmajor(ah + o, CH), desc_mnmajor(b_hi + o, CH), idesc, 1u);
        }
        mma_commit(&bar[BB_FFULL + buf]);
        piece++;
      };
      for (int n = 0; n < L; n++) {
        const uint32_t ph = n & 1;
        wait(BB_GDP, ph);                                     // G, dpre, LN1(h) (and d) are in place
        for (int f = 0; f < 3; f++)                           // dW1 += G^T d            fullconnect.py:118-121
          for (int j = 0; j < 2; j++) issue(caddr(C_GH), caddr(C_GL), f, caddr(C_DH + j), caddr(C_DL + j), j ? 32 : 64);
        mma_commit(&bar[BB_DW1_DONE]);
        for (int f = 0; f < 2; f++)                           // dW0 (+ db0) += [LN1(h) | 1]^T dpre
          for (int j = 0; j < 3; j++) issue(caddr(C_U1H), caddr(C_U1L), f, caddr(C_PH + j), caddr(C_PL + j), 64);
        mma_commit(&bar[BB_DW0_DONE]);
        wait(BB_DQKV, ph);
        wait(BB_U, ph);
        for (int f = 0; f < 2; f++)                           // dWqkv (+ dbqkv) += [LN(x) | 1]^T dqkv
          for (int j = 0; j < 6; j++) issue(caddr(C_UH), caddr(C_UL), f, caddr(C_QH + j), caddr(C_QL + j), 64);
        mma_commit(&bar[BB_DWQKV_DONE]);
      }
    }
    __syncwarp();
  } else if (warp == 15) {
    // ===================== column sums (bias / LayerNorm gradients): shared accumulators -> red.global.add =====================
    for (int n = 0; n < L; n++) {
      const NtVitBlock& P = a.blk[L - 1 - n];
      mbar_wait(&bar[BB_ACCRDY], n & 1);
      float* acc = accs;
      float* dsts[5] = {P.d_b1, P.d_ln1_g, P.d_ln1_b, P.d_ln_g, P.d_ln_b};
#pragma unroll
      for (int k = 0; k < 5; k++)
        for (int c = lane; c < D; c += 32) {
          atomicAdd(dsts[k] + c, acc[96 * k + c]);
          acc[96 * k + c] = 0.f;
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[BB_ACCFREE]);
    }
  } else if (warp == 6 || warp == 7 || warp == 10 || warp == 11) {
    // ===================== weight-gradient flush: TMEM lanes 64..127 -> shared staging -> red.global.add.v4 ================
    // A lane holds 32 columns of ONE gradient row; straight from registers every red instruction would touch 32 rows (32
    // half-used sectors) and the flood fills this SM's load/store queue in front of the token warps' shuffles and shared
    // memory traffic (their LayerNorm-backward phase took 20 K cycles for 400 instructions).  Through a small staging
    // tile each instruction instead covers 8 rows x 64 contiguous bytes.  (cp.reduce.async.bulk, one per row, was tried:
    // the TMA engine needs ~20 cycles per 128-byte operation, 3072 of them per block — slower than the block itself.)
    const int fq = warp & 3, fc = (warp >> 2) - 1;                     // TMEM lane quadrant (2, 3); 32-column half of the piece
    uint8_t* tile = smem + BOFF_STG + (uint32_t)(2 * fc + fq - 2) * 32 * STG_PITCH;
    const int sub = lane >> 2, quad = lane & 3;                        // read side: row 8 i + sub, columns 4 quad .. 4 quad + 3
    const int F0 = 32 * (fq - 2);                                      // first feature (within the 64-feature half) of this warp
    uint32_t piece = 0;
    for (int n = 0; n < L; n++) {
      const NtVitBlock& P = a.blk[L - 1 - n];
      for (int p = 0; p < PIECES; p++, piece++) {
        const int buf = piece & 1;
        mbar_wait(&bar[BB_FFULL + buf], (piece >> 1) & 1);
        tc_fence_after();
        uint32_t v[32];
        tmem_ld32(tbase + ((uint32_t)(32 * fq) << 16) + 384 + 64 * buf + 32 * fc, v);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[BB_FFREE + buf]);               // accumulator consumed: the MMA issuer may reuse it
        // destination of staged row 0 of this warp, row pitch, valid rows [0, nrow) plus an optional bias row
        float* base = nullptr;
        float* bias = nullptr;
        int pitch = 0, nrow = 32, brow = -1, ncol = 32;
        if (p < 6) {
          const int f = p >> 1, j = p & 1;
          base = P.d_w1 + (size_t)(64 * f + F0) * D + 64 * j + 32 * fc; pitch = D;
          if (j == 1 && fc == 1) ncol = 0;                              // the second column piece of dW1 is 32 wide
        } else if (p < 12) {
          const int f = (p - 6) / 3, j = (p - 6) % 3, Fa = 64 * f + F0;
          base = P.d_w0 + (size_t)Fa * HD + 64 * j + 32 * fc; pitch = HD;
          nrow = min(32, max(0, D - Fa));
          if (D >= Fa && D < Fa + 32) { brow = D - Fa; bias = P.d_b0 + 64 * j + 32 * fc; }
        } else {
          const int f = (p - 12) / 6, j = (p - 12) % 6, Fa = 64 * f + F0;
          const int col = (j >> 1) * D + DH * (2 * (j & 1) + fc);       // [q|k|v][head][24] <- padded [part][head][32]
          base = P.d_wqkv + (size_t)Fa * 3 * D + col; pitch = 3 * D;
          nrow = min(32, max(0, D - Fa));
          if (D >= Fa && D < Fa + 32) { brow = D - Fa; bias = P.d_bqkv + col; }
          ncol = DH;
        }
#pragma unroll
        for (int hc = 0; hc < 2; hc++) {                                // two 16-column halves through the tile
          __syncwarp();                                                 // the previous half's reads of the tile are done
#pragma unroll
          for (int c = 0; c < 16; c += 4)
            *reinterpret_cast<uint4*>(tile + lane * STG_PITCH + 4 * c) = make_uint4(v[16 * hc + c], v[16 * hc + c + 1], v[16 * hc + c + 2], v[16 * hc + c + 3]);
          __syncwarp();
          if (16 * hc + 4 * quad < ncol) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const int row = 8 * i + sub;
              const float4 t = *reinterpret_cast<const float4*>(tile + row * STG_PITCH + 16 * quad);
              if (row < nrow) red4(base + (size_t)row * pitch + 16 * hc + 4 * quad, t.x, t.y, t.z, t.w);
              else if (row == brow) red4(bias + 16 * hc + 4 * quad, t.x, t.y, t.z, t.w);
            }
          }
        }
      }
    }
  } else if (is_tok) {
    // ===================== token warps: the backward chain of one token row =====================
    const bool live = r < N;
    const float scale = rsqrtf((float)DH);
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
    auto arrive = [&](int b) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[b]);
    };
    auto wait = [&](int b, uint32_t ph) { mbar_wait(&bar[b], ph); tc_fence_after(); };
    float d[24];                                             // gradient w.r.t. the current block's output
    if (live) load24_global(a.dout + ((size_t)img * N + r) * D + 24 * g, d);
    else {
#pragma unroll
      for (int c = 0; c < 24; c++) d[c] = 0.f;
    }
    float* dh_scratch = a.dx + (size_t)img * N * D;           // dh parked here (tiled) during the attention phase
    for (int n = 0; n < L; n++) {
      const int bi = L - 1 - n;
      const NtVitBlock& P = a.blk[bi];
      const uint32_t ph = n & 1;
      float v[24], w[24];
      const bool t0 = tid == 0;
      if (t0) stamp(a, 0, n, 0);
      float* acc = accs + 24 * g;                             // [quantity][96]: db1, dgamma1, dbeta1, dgamma, dbeta of this block
      const float* pv = par + (n & 1) * BPAR_FLOATS + 24 * g;  // this thread's 24 columns of ln_g | ln_b | ln1_g | ln1_b (stride 96)
      mbar_wait(&bar[BB_PAR + (n & 1)], (n >> 1) & 1);
      // ---- d -> operand planes
      if (n > 0 && !(a.dbg_flags & 1)) wait(BB_DW1_DONE, (n - 1) & 1);     // the previous block's dW1 products read the d planes
      store24_split(chunk(C_DH), chunk(C_DL), r, 24 * g, d);
      arrive(BB_D);
      if (t0) stamp(a, 0, n, 1);
      // ---- u1 = LN1(h) (recomputed), xhat1 kept for the LayerNorm backward
      float xh[24], rstd1;
      {
        if (live) load24_tiled(P.h + (size_t)img * N * D, N, r, 6 * g, v);
        else {
#pragma unroll
          for (int c = 0; c < 24; c++) v[c] = 0.f;
        }
        float mean;
        row_stats(v, stat, r, g, mean, rstd1);
#pragma unroll
        for (int c = 0; c < 24; c++) {
          xh[c] = live ? (v[c] - mean) * rstd1 : 0.f;
          w[c] = live ? xh[c] * pv[192 + c] + pv[288 + c] : 0.f;
        }
        if (n > 0 && !(a.dbg_flags & 1)) wait(BB_DWQKV_DONE, (n - 1) & 1);   // the previous block's dWqkv products read this area (LN(x), dqkv)
        store24_split(chunk(C_U1H), chunk(C_U1L), r, 24 * g, w);
        if (g == 0) store_ones_unit(chunk(C_U1H), chunk(C_U1L), r, D, live);
        // db1 += colsum(d)  (fullconnect.py:122) — here, not at the top: the previous block's column sums have left the
        // shared accumulators by now
        if (n > 0) wait(BB_ACCFREE, (n - 1) & 1);
        colsum24_atomic(d, lane, acc);
      }
      // ---- dG -> G = GELU(pre), dpre = dG * GELU'(pre)                   attention.py:58-59
      if (t0) stamp(a, 0, n, 2);
      {
        const float* pre_img = P.dgelu + (size_t)img * N * HD;
#pragma unroll 1
        for (int it = 0; it < 6; it++) {                       // rolled over 8-column units (code size: see the forward)
          const int nt = it / 3, j = it - 3 * nt;
          if (j == 0) {
            wait(BB_DGACC + nt, ph);
            if (t0) stamp(a, 0, n, 3 + nt);
          }
          uint32_t t[8];
          tmem_ld8(lane_addr + NT * nt + 24 * g + 8 * j, t);
          float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
          if (live) {
            const int u = 24 * nt + 6 * g + 2 * j;
            p0 = *reinterpret_cast<const float4*>(pre_img + ((size_t)u * N + r) * 4);
            p1 = *reinterpret_cast<const float4*>(pre_img + ((size_t)(u + 1) * N + r) * 4);
          }
          tmem_wait_ld();
          const float pre[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
          float gy[8], dp[8];
#pragma unroll
          for (int i = 0; i < 8; i++) {
            float y, dy;
            gelu_both(pre[i], y, dy);
            gy[i] = live ? y : 0.f;
            dp[i] = live ? __uint_as_float(t[i]) * dy : 0.f;
          }
          store8_split(chunk(C_GH), chunk(C_GL), r, NT * nt + 24 * g + 8 * j, CH, gy);
          store8_split(chunk(C_PH), chunk(C_PL), r, NT * nt + 24 * g + 8 * j, CH, dp);
        }
      }
      arrive(BB_GDP);
      if (t0) stamp(a, 0, n, 5);
      // ---- dh = d + LN1-backward(dU1); dgamma1, dbeta1                    attention.py:61-62, layernorm.py:116-135
      {
        wait(BB_DU1ACC, ph);
        if (t0) stamp(a, 0, n, 6);
        ld24(lane_addr + 192 + 24 * g, v);
        if (t0) stamp(a, 1, n, 0);
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int c = 0; c < 24; c++) {
          if (!live) v[c] = 0.f;
          w[c] = v[c] * xh[c];                                 // dgamma contribution
          const float gd = v[c] * pv[192 + c];
          p1 += gd;
          p2 += gd * xh[c];
        }
        if (t0) stamp(a, 1, n, 1);
        colsum24_atomic(w, lane, acc + 96);
        colsum24_atomic(v, lane, acc + 192);
        if (t0) stamp(a, 1, n, 2);
        float s1, s2;
        row_sum2(p1, p2, stat, r, g, s1, s2);
        if (t0) stamp(a, 1, n, 3);
        s1 *= (1.0f / D); s2 *= (1.0f / D);
#pragma unroll
        for (int c = 0; c < 24; c++) {
          const float gd = v[c] * pv[192 + c];
          d[c] = live ? d[c] + rstd1 * (gd - s1 - xh[c] * s2) : 0.f;
        }
        if (t0) stamp(a, 1, n, 4);
        if (live) store24_tiled(dh_scratch, N, r, 6 * g, d);
        if (t0) stamp(a, 1, n, 5);
      }
      // ---- dO (= dh) and q, k, v -> head-padded planes
      {
        if (t0) stamp(a, 0, n, 7);
        if (!(a.dbg_flags & 1)) wait(BB_DW0_DONE, ph);         // dW1 / dW0 products have read d, G, dpre, LN1(h)
        if (t0) stamp(a, 0, n, 8);
        store24_split(chunk(C_DH), chunk(C_DL), r, DHP * g, d);
        {
          const uint32_t zoff = plane_off(r, DHP * g + DH, CH);
          *reinterpret_cast<uint4*>(chunk(C_DH) + zoff) = make_uint4(0, 0, 0, 0);
          *reinterpret_cast<uint4*>(chunk(C_DL) + zoff) = make_uint4(0, 0, 0, 0);
        }
        {
          const float* qkv_img = P.qkv + (size_t)img * N * 3 * D;
#pragma unroll 1
          for (int it = 0; it < 9; it++) {
            const int part = it / 3, j = it - 3 * part;
            float4 p0 = make_float4(0.f, 0.f, 0.f, 0.f), p1 = p0;
            if (live) {
              const int u = 24 * part + 6 * g + 2 * j;
              p0 = *reinterpret_cast<const float4*>(qkv_img + ((size_t)u * N + r) * 4);
              p1 = *reinterpret_cast<const float4*>(qkv_img + ((size_t)(u + 1) * N + r) * 4);
            }
            const float w8[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
            uint8_t* hi = chunk(C_QH + 2 * part);
            uint8_t* lo = chunk(C_QL + 2 * part);
            store8_split(hi, lo, r, DHP * g + 8 * j, CH, w8);
            if (j == 2) {
              const uint32_t zoff = plane_off(r, DHP * g + DH, CH);
              *reinterpret_cast<uint4*>(hi + zoff) = make_uint4(0, 0, 0, 0);
              *reinterpret_cast<uint4*>(lo + zoff) = make_uint4(0, 0, 0, 0);
            }
          }
        }
        arrive(BB_ATT);
        if (t0) stamp(a, 0, n, 9);
      }
      // ---- head g: P (recomputed), dS = P o (dP - rowsum(dP o P)) / sqrt(dh)        attention.py:74-75
      {
        uint32_t s0[32], s1[32];
        wait(BB_SACC + g, ph);
        if (t0) stamp(a, 0, n, 10);
        tmem_ld32(lane_addr + 64 * g, s0);
        tmem_ld32(lane_addr + 64 * g + 32, s1);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[BB_SCONS + g]);
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
#pragma unroll
        for (int j = 0; j < 32; j++) {
          s0[j] = __float_as_uint(__uint_as_float(s0[j]) * inv);
          s1[j] = __float_as_uint(__uint_as_float(s1[j]) * inv);
        }
        uint8_t* ps_hi = chunk(ps_chunk(g));
        uint8_t* ps_lo = ps_hi + CH;
#pragma unroll
        for (int j8 = 0; j8 < 8; j8++) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; i++) t[i] = __uint_as_float(8 * j8 + i < 32 ? s0[(8 * j8 + i) & 31] : s1[(8 * j8 + i) & 31]);
          store8_split(ps_hi, ps_lo, r, 8 * j8, CH, t);
        }
        arrive(BB_P + g);
        if (t0) stamp(a, 0, n, 11);
        // row sum of dP o P (dP is read from TMEM twice; P stays in registers)
        wait(BB_DPACC + g, ph);
        float tsum = 0.f;
        {
          uint32_t dp[32];
          tmem_ld32(lane_addr + 64 * g, dp);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) tsum += __uint_as_float(dp[j]) * __uint_as_float(s0[j]);
          tmem_ld32(lane_addr + 64 * g + 32, dp);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) tsum += __uint_as_float(dp[j]) * __uint_as_float(s1[j]);
          tmem_ld32(lane_addr + 64 * g, dp);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) s0[j] = __float_as_uint(__uint_as_float(s0[j]) * (__uint_as_float(dp[j]) - tsum) * scale);
          tmem_ld32(lane_addr + 64 * g + 32, dp);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 32; j++) s1[j] = __float_as_uint(__uint_as_float(s1[j]) * (__uint_as_float(dp[j]) - tsum) * scale);
        }
        wait(BB_DVACC + g, ph);                                // dV_h has read P: the buffer takes dS
#pragma unroll
        for (int j8 = 0; j8 < 8; j8++) {
          float t[8];
#pragma unroll
          for (int i = 0; i < 8; i++) t[i] = live ? __uint_as_float(8 * j8 + i < 32 ? s0[(8 * j8 + i) & 31] : s1[(8 * j8 + i) & 31]) : 0.f;
          store8_split(ps_hi, ps_lo, r, 8 * j8, CH, t);
        }
        arrive(BB_DS + g);
        if (t0) stamp(a, 0, n, 12);
      }
      // ---- dq, dk, dv of head g -> the planes q, k, v occupied (attention.py:81-83)
      {
        wait(BB_DQKACC + g, ph);
        if (t0) stamp(a, 0, n, 13);
#pragma unroll 1
        for (int it = 0; it < 9; it++) {
          const int part = it / 3, j = it - 3 * part;
          const uint32_t col = (part == 0 ? 64 * g : (part == 1 ? 64 * g + 32 : 256 + DHP * g)) + 8 * j;
          uint32_t t[8];
          tmem_ld8(lane_addr + col, t);
          tmem_wait_ld();
          float w8[8];
#pragma unroll
          for (int i = 0; i < 8; i++) w8[i] = live ? __uint_as_float(t[i]) : 0.f;
          store8_split(chunk(C_QH + 2 * part), chunk(C_QL + 2 * part), r, DHP * g + 8 * j, CH, w8);
        }
        arrive(BB_DQKV);
      }
      // ---- u = LN(x) (recomputed) for dWqkv; xhat kept for the LayerNorm backward
      float rstd0;
      {
        const float* xin = bi ? a.blk[bi - 1].out : a.x;
        if (live) {
          if (bi) load24_tiled(xin + (size_t)img * N * D, N, r, 6 * g, v);
          else load24_global(xin + ((size_t)img * N + r) * D + 24 * g, v);
        } else {
#pragma unroll
          for (int c = 0; c < 24; c++) v[c] = 0.f;
        }
        float mean;
        row_stats(v, stat, r, g, mean, rstd0);
#pragma unroll
        for (int c = 0; c < 24; c++) {
          xh[c] = live ? (v[c] - mean) * rstd0 : 0.f;
          w[c] = live ? xh[c] * pv[c] + pv[96 + c] : 0.f;
        }
        wait(BB_DQKACC + 1, ph);                               // heads 0, 1 kept P / dS here; their products are done
        store24_split(chunk(C_UH), chunk(C_UL), r, 24 * g, w);
        if (g == 0) store_ones_unit(chunk(C_UH), chunk(C_UL), r, D, live);
        arrive(BB_U);
        if (t0) stamp(a, 0, n, 14);
      }
      // ---- dx = dh + LN-backward(dU); dgamma, dbeta                       attention.py:87-88
      {
        wait(BB_DUACC, ph);
        if (t0) stamp(a, 0, n, 15);
        ld24(lane_addr + 24 * g, v);
        float p1 = 0.f, p2 = 0.f;
#pragma unroll
        for (int c = 0; c < 24; c++) {
          if (!live) v[c] = 0.f;
          w[c] = v[c] * xh[c];
          const float gd = v[c] * pv[c];
          p1 += gd;
          p2 += gd * xh[c];
        }
        colsum24_atomic(w, lane, acc + 288);
        colsum24_atomic(v, lane, acc + 384);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bar[BB_ACCRDY]);           // (mbarrier arrive has release semantics for the shared atomics)
        // dh back from its parking place BEFORE the row-sum barriers: after the last block the same buffer receives dx in
        // row-major order, and no thread may write that while another still has to read its parked dh
        if (live) load24_tiled(dh_scratch, N, r, 6 * g, d);
        float s1, s2;
        row_sum2(p1, p2, stat, r, g, s1, s2);
        s1 *= (1.0f / D); s2 *= (1.0f / D);
#pragma unroll
        for (int c = 0; c < 24; c++) {
          const float gd = v[c] * pv[c];
          d[c] = live ? d[c] + rstd0 * (gd - s1 - xh[c] * s2) : 0.f;
        }
      }
    }
    if (live) store24_global(a.dx + ((size_t)img * N + r) * D + 24 * g, d);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    tmem_dealloc<512>(tbase);
  }
}

static bool supported(int N, int D_, int H_) { return N >= 1 && N <= ROWS && D_ == D && H_ == H; }
static long long* g_dbg = nullptr;

}  // namespace vtc
}  // namespace nt
using namespace nt;

extern "C" {

int nt_vit_tc_supported(int N, int D, int H, int* ok) {
  *ok = vtc::supported(N, D, H) ? 1 : 0;
  return 0;
}

/* development aid: phase time stamps of CTA 0 ([2 roles][16 blocks][16 events] clock64 values); NULL switches them off */
int nt_vit_tc_debug(long long* stamps) {
  vtc::g_dbg = stamps;
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

int nt_vit_tc_bwd(const NtVitBlock* blocks, int nblocks, const float* x, const float* dout, float* dx, int B, int N, int D, int H) {
  NT_ENTRY();
  vtc::Args a{};
  if (int rc = vtc_fill(a, blocks, nblocks, "nt_vit_tc_bwd", N, D, H)) return rc;
  if (B == 0) return 0;
  a.x = x; a.dout = dout; a.dx = dx;
  a.dbg = vtc::g_dbg;
  if (const char* e = getenv("NTB200_VTC_DBG")) a.dbg_flags = atoi(e);
  static bool attr = false;
  if (!attr) {
    NT_CHECK(cudaFuncSetAttribute(vtc::vit_tc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vtc::BSMEM_BYTES));
    attr = true;
  }
  NT_CHECK(launch_k(vtc::vit_tc_bwd_kernel, dim3(B), dim3(vtc::BTHREADS), vtc::BSMEM_BYTES, a));
  NT_LAUNCH_OK();
  return 0;
}

int nt_vit_tc_fwd(const NtVitBlock* blocks, int nblocks, const float* x, int B, int N, int D, int H, int prepared) {
  NT_ENTRY();
  vtc::Args a{};
  if (int rc = vtc_fill(a, blocks, nblocks, "nt_vit_tc_fwd", N, D, H)) return rc;
  if (B == 0) return 0;
  a.x = x;
  a.dbg = vtc::g_dbg;
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
