// Attention forward on the 5th-generation tensor cores for dh = 64 and 64 < N <= 256 tokens (the GPT configs: T = 256,
// causal): attention.py:31-46 / gpt/attdecoderblock.py:52-69.  One work item = (sample, head, tile of 128 query rows);
// persistent CTAs walk the items, heavy (late, causal) tiles first.
//
// Construction as in attention_tc.cu: q, k, v and P live in shared memory as bf16 hi / lo planes in the 128-byte-swizzled
// UMMA layout; S = Q K^T (M = 128, N = keys <= 256, K = 64) and O = P V (M = 128, N = 64, K = keys) are three-pass
// tcgen05.mma products (lo*hi + hi*lo + hi*hi, fp32 accumulators in TMEM) issued by one thread.  Accumulator row i is
// TMEM lane i; with 128 query rows all four lane quadrants carry rows, so the softmax runs on all four SM sub-partitions:
// 16 token warps, thread = (row, quarter of the key columns), the four threads of a row exchange their partial maximum /
// sum through shared memory.  The whole row is in TMEM, so there is no online rescaling.  P overwrites the q / k planes
// (dead once S is complete); causal tiles only load and multiply the keys they can see.
//
// Two operand paths.  PLANES: the QKV Linear's epilogue already wrote qkv as row-major bf16 hi / lo planes
// (ops.linear_fwd(split_out=True)); one thread lands q, k, v tiles straight into the swizzled operand layout with TMA
// (3-D tensor map {3D, N, B}, box {64, 128, 1}: rows beyond the sequence are zero-filled) one item ahead, so the token
// warps never touch the operands.  Otherwise the token warps read the fp32 tensor and split it themselves.
#include "common.cuh"
#include "umma.cuh"
#include <cuda.h>
#include <cstdlib>
#include <cstring>

namespace nt {
namespace atl {
using namespace um;

constexpr int DH = 64, BM = 128, MAXK = 256;
constexpr uint32_t PCH = BM * 128;              // one 64-column chunk of a 128-row plane (16 KB)
constexpr int TOK_WARPS = 16, TOK_THREADS = 32 * TOK_WARPS;
constexpr int THREADS = TOK_THREADS;            // thread 0 also issues the MMAs (the chain is serial per item; 512 threads
                                                // leave 128 registers per thread for the 64 scores a thread holds)

// shared memory map.  P (hi: 4 chunks, lo: 4 chunks) aliases q and k, which are dead when S is complete.
constexpr uint32_t OFF_P_HI = 0, OFF_P_LO = 4 * PCH;
constexpr uint32_t OFF_Q_HI = 0, OFF_Q_LO = PCH, OFF_K_HI = 2 * PCH, OFF_K_LO = 4 * PCH;      // k planes: [256][128 B] = 32 KB
constexpr uint32_t OFF_V_HI = 8 * PCH, OFF_V_LO = 10 * PCH;
constexpr uint32_t OFF_OST = 6 * PCH;           // fp32 staging tile of O [128][64]: the tail of P lo (dead once O is complete)
constexpr uint32_t OFF_XCH = 12 * PCH;          // [128 rows][4 quarters][max, sum]
constexpr uint32_t OFF_BAR = OFF_XCH + BM * 4 * 2 * 4;
enum { B_QK = 0, B_V, B_S, B_P, B_O, B_COUNT };
constexpr uint32_t SMEM = OFF_BAR + B_COUNT * 8 + 16;
static_assert(SMEM <= 232448, "shared memory budget");
constexpr uint32_t T_S = 0, T_O = 256;          // TMEM columns

struct Args {
  const float* qkv;      // [B,N,3*H*64]
  float* out;            // [B,N,H*64]
  float* P;              // [B,H,N,N], or null: not materialised (the tcgen05 backward recomputes it from lse)
  const float* residual; // [B,N,H*64] or null
  uint32_t* o_hi;        // bf16 pairs of `out` (row-major) or null
  uint32_t* o_lo;
  float* lse;            // [B,H,N] or null: log2-domain log-sum-exp of every score row, max * c + log2(sum), c = log2(e) / sqrt(dh)
                         // (what the tcgen05 backward needs to recompute P: P_ij = 2^(s_ij c - lse_i))
  int B, N, H, causal;
  int planes;            // q, k, v arrive by TMA from the bf16 planes of qkv (tensor maps in the kernel parameters)
  long long* dbg;        // optional phase time stamps of CTA 0: [item < 8][warp 0 / warp 15][16 events]
};
static long long* g_dbg = nullptr;
static long long* g_dbg_bwd = nullptr;
__device__ __forceinline__ void stamp(const Args& a, int n, int ev) {
  if (a.dbg && blockIdx.x == 0 && n < 8 && (threadIdx.x == 0 || threadIdx.x == 480)) a.dbg[(n * 2 + (threadIdx.x ? 1 : 0)) * 16 + ev] = clock64();
}

// rows [row0, row0 + nrows) of one of q / k / v of head h (64 columns, row pitch 3D floats) -> a hi / lo plane pair.
// Unit = (row, 8 columns): a warp reads 4 rows x 256 contiguous bytes per pass.  Loads and plane stores are separate so
// that the loads of q, k and v are all in flight before the first store.
template <int U>
__device__ __forceinline__ void ld_units(const float* __restrict__ src, int row0, int nrows, int N, int pitch, int tid, float4 (&t)[U][2]) {
#pragma unroll
  for (int j = 0; j < U; j++) {
    const int u = tid + TOK_THREADS * j, row = u >> 3, c8 = u & 7;
    const bool live = row < nrows && row0 + row < N;
    const float4* p = reinterpret_cast<const float4*>(src + (long long)(row0 + row) * pitch + 8 * c8);
    t[j][0] = live ? __ldg(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    t[j][1] = live ? __ldg(p + 1) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}
template <int U>
__device__ __forceinline__ void st_units(uint8_t* hi, uint8_t* lo, int nrows, int tid, const float4 (&t)[U][2]) {
#pragma unroll
  for (int j = 0; j < U; j++) {
    const int u = tid + TOK_THREADS * j, row = u >> 3, c8 = u & 7;
    if (row < nrows) {
      const float v[8] = {t[j][0].x, t[j][0].y, t[j][0].z, t[j][0].w, t[j][1].x, t[j][1].y, t[j][1].z, t[j][1].w};
      store8_split(hi, lo, row, 8 * c8, 0, v);
    }
  }
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const CUtensorMap* tm, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool PLANES>
__global__ void __launch_bounds__(THREADS, 1) attn_tcl_fwd_kernel(Args a, const __grid_constant__ CUtensorMap tm_hi,
                                                                  const __grid_constant__ CUtensorMap tm_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  float* xch = reinterpret_cast<float*>(smem + OFF_XCH);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + B_COUNT);

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, H = a.H, D = H * DH;
  const int QT = (N + BM - 1) / BM, BH = a.B * H, items = BH * QT;
  const int Npad = (N + 63) & ~63;

  if (tid == 0) {
    mbar_init(&bar[B_QK], PLANES ? 1 : TOK_WARPS);
    mbar_init(&bar[B_V], PLANES ? 1 : TOK_WARPS);
    mbar_init(&bar[B_P], TOK_WARPS);
    mbar_init(&bar[B_S], 1);
    mbar_init(&bar[B_O], 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  pdl_wait();

  // item -> (sample*head, query tile); causal: the late tiles (more keys) first, so that the tail is made of light items
  auto decode = [&](int it, int& bh, int& qt, int& nk) {
    qt = QT - 1 - it / BH;
    bh = it % BH;
    nk = a.causal ? min(Npad, (qt + 1) * BM) : Npad;
  };

  const uint32_t sb = smem_u32(smem);
  // PLANES: thread 0 lands q, k (barrier B_QK) and v (B_V) of item `it` in the operand planes
  auto tma_item = [&](int it) {
    int bh, qt, nk;
    decode(it, bh, qt, nk);
    const int b = bh / H, c = (bh % H) * DH;
    const int kb = (nk + 127) >> 7;                                  // 128-row boxes of keys
    mbar_expect_tx(&bar[B_QK], 2u * PCH + 2u * (uint32_t)kb * PCH);
    tma_load_3d(sb + OFF_Q_HI, &tm_hi, &bar[B_QK], c, qt * BM, b);
    tma_load_3d(sb + OFF_Q_LO, &tm_lo, &bar[B_QK], c, qt * BM, b);
    for (int j = 0; j < kb; j++) {
      tma_load_3d(sb + OFF_K_HI + j * PCH, &tm_hi, &bar[B_QK], D + c, j * 128, b);
      tma_load_3d(sb + OFF_K_LO + j * PCH, &tm_lo, &bar[B_QK], D + c, j * 128, b);
    }
    mbar_expect_tx(&bar[B_V], 2u * (uint32_t)kb * PCH);
    for (int j = 0; j < kb; j++) {
      tma_load_3d(sb + OFF_V_HI + j * PCH, &tm_hi, &bar[B_V], 2 * D + c, j * 128, b);
      tma_load_3d(sb + OFF_V_LO + j * PCH, &tm_lo, &bar[B_V], 2 * D + c, j * 128, b);
    }
  };
  // S = Q K^T of the current item (thread 0, once every warp has delivered its share of the q / k planes)
  auto issue_s = [&](int nk, uint32_t ph) {
    const uint32_t id_s = idesc_bf16(128, nk, 0, 0);
    mbar_wait(&bar[B_QK], ph);
    tc_fence_after();
#pragma unroll
    for (int ks = 0; ks < 4; ks++) {
      const uint64_t ah = desc_kmajor(sb + OFF_Q_HI + ks * KMAJ_STEP), al = desc_kmajor(sb + OFF_Q_LO + ks * KMAJ_STEP);
      const uint64_t bh_ = desc_kmajor(sb + OFF_K_HI + ks * KMAJ_STEP), bl = desc_kmajor(sb + OFF_K_LO + ks * KMAJ_STEP);
      mma_bf16(tbase + T_S, al, bh_, id_s, ks ? 1u : 0u);
      mma_bf16(tbase + T_S, ah, bl, id_s, 1u);
      mma_bf16(tbase + T_S, ah, bh_, id_s, 1u);
    }
    mma_commit(&bar[B_S]);
  };
  // O = P V.  Per K = 16 step two MMAs instead of three: P.hi x [V.hi | V.lo] (the two planes are one MN-major operand
  // of N = 128, chunk stride = the distance between the planes) lands hi*hi in accumulator columns 0..63 and hi*lo in
  // 64..127, P.lo x V.hi adds to 0..63; the epilogue sums the halves.  (An M = 128, K = 16 MMA has a floor of ~51 cycles
  // however narrow it is — tools/umma_rate.cu: N = 64 51, N = 128 66 — so fewer, wider MMAs are cheaper.)
  auto issue_o = [&](int nk, uint32_t ph) {
    const uint32_t id_o = idesc_bf16(128, 64, 0, 1), id_o2 = idesc_bf16(128, 128, 0, 1);
    const uint64_t ah0 = desc_kmajor(sb + OFF_P_HI), al0 = desc_kmajor(sb + OFF_P_LO);
    const uint64_t bw0 = desc_mnmajor(sb + OFF_V_HI, OFF_V_LO - OFF_V_HI), bh0 = desc_mnmajor(sb + OFF_V_HI, 8192);
    mbar_wait(&bar[B_V], ph);
    mbar_wait(&bar[B_P], ph);
    tc_fence_after();
    const int ksteps = nk >> 4;
#pragma unroll 4
    for (int ks = 0; ks < ksteps; ks++) {
      // the start-address field counts 16-byte units: advancing an operand is an add on the descriptor
      const uint64_t pa = (uint64_t)((ks >> 2) * (PCH >> 4) + (ks & 3) * (KMAJ_STEP >> 4)), pb = (uint64_t)(ks * (MNMAJ_STEP >> 4));
      mma_bf16(tbase + T_O, ah0 + pa, bw0 + pb, id_o2, ks ? 1u : 0u);
      mma_bf16(tbase + T_O, al0 + pa, bh0 + pb, id_o, 1u);
    }
    mma_commit(&bar[B_O]);
  };
  {
    // ===================== token warps: thread = (query row r, quarter cg of the key columns) =====================
    const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;
    const float scale = rsqrtf((float)DH);
    const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
    auto arrive = [&](int b) {
      tc_fence_before();
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bar[b]);
    };
    uint32_t n = 0;
    // q and k of an item are fetched into registers one item ahead (under the P V product and the output epilogue of the
    // previous item, when the 64 score registers are free); v is fetched at the top of the item, under the plane stores
    float4 tq[2][2], tk[4][2];
    auto fetch_qk = [&](int it) {
      int bh, qt, nk;
      decode(it, bh, qt, nk);
      const float* base = a.qkv + (long long)(bh / H) * N * 3 * D + (bh % H) * DH;
      ld_units<2>(base, qt * BM, BM, N, 3 * D, tid, tq);
      ld_units<4>(base + D, 0, nk, N, 3 * D, tid, tk);
    };
    if ((int)blockIdx.x < items) {
      if constexpr (!PLANES) fetch_qk(blockIdx.x);
      else if (tid == 0) tma_item(blockIdx.x);
    }
    for (int it = blockIdx.x; it < items; it += gridDim.x, n++) {
      int bh, qt, nk;
      decode(it, bh, qt, nk);
      const uint32_t ph = n & 1;
      const int b = bh / H, h = bh % H;
      const int grow = qt * BM + r;                       // this thread's query row inside the sequence
      const bool live = grow < N;
      const float* base = a.qkv + (long long)b * N * 3 * D + h * DH;
      stamp(a, n, 0);
      // ---- operands -> planes (q: 128 rows, k / v: the nk visible keys); v is fetched under the S product
      if constexpr (!PLANES) {
        st_units<2>(smem + OFF_Q_HI, smem + OFF_Q_LO, BM, tid, tq);
        st_units<4>(smem + OFF_K_HI, smem + OFF_K_LO, nk, tid, tk);
        stamp(a, n, 1);
        arrive(B_QK);
        float4 tv[4][2];
        ld_units<4>(base + 2 * D, 0, nk, N, 3 * D, tid, tv);
        if (tid == 0) issue_s(nk, ph);
        __syncwarp();
        st_units<4>(smem + OFF_V_HI, smem + OFF_V_LO, nk, tid, tv);
        stamp(a, n, 2);
        arrive(B_V);
      } else {
        if (tid == 0) {
          if (it + (int)gridDim.x < items) {
            // the next item's tiles towards L2 now: a tile is 128-byte rows 3D * 2 bytes apart, slow when it has to come from
            // DRAM at the moment it is needed
            int bh2, qt2, nk2;
            decode(it + gridDim.x, bh2, qt2, nk2);
            const int b2 = bh2 / H, c2 = (bh2 % H) * DH;
            tma_prefetch_3d(&tm_hi, c2, qt2 * BM, b2);
            tma_prefetch_3d(&tm_lo, c2, qt2 * BM, b2);
            for (int j = 0; j < ((nk2 + 127) >> 7); j++) {
              tma_prefetch_3d(&tm_hi, D + c2, j * 128, b2);
              tma_prefetch_3d(&tm_lo, D + c2, j * 128, b2);
              tma_prefetch_3d(&tm_hi, 2 * D + c2, j * 128, b2);
              tma_prefetch_3d(&tm_lo, 2 * D + c2, j * 128, b2);
            }
          }
          issue_s(nk, ph);
        }
        __syncwarp();
      }
      stamp(a, n, 3);
      // ---- P = softmax(S / sqrt(dh) (+ causal mask)); this thread's nk / 4 key columns           attention.py:37-41
      const int cpt = nk >> 2, col0 = cg * cpt, n16 = cpt >> 4;        // 16, 32, 48 or 64 columns per thread
      uint32_t p[64];
      mbar_wait(&bar[B_S], ph);
      tc_fence_after();
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < n16) tmem_ld16(lane_addr + T_S + col0 + 16 * i, reinterpret_cast<uint32_t(&)[16]>(p[16 * i]));
      tmem_wait_ld();
      stamp(a, n, 4);
      // row maximum of the RAW scores (the scale is positive); masked entries become -inf and stay out of everything after.
      // Only the blocks of 16 columns that touch the diagonal / the ragged end test per element (warp-uniform choice).
      float mx = -INFINITY;
      const int rowmin = qt * BM + 32 * q;
#pragma unroll
      for (int i = 0; i < 4; i++)
        if (i < n16) {
          const int cb = col0 + 16 * i;
          if (cb + 15 >= N || (a.causal && cb + 15 > rowmin)) {
#pragma unroll
            for (int j = 0; j < 16; j++) {
              const int col = cb + j;
              const bool ok = col < N && !(a.causal && col > grow);
              const float sv = ok ? __uint_as_float(p[16 * i + j]) : -INFINITY;
              p[16 * i + j] = __float_as_uint(sv);
              mx = fmaxf(mx, sv);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++) mx = fmaxf(mx, __uint_as_float(p[16 * i + j]));
          }
        }
      float2* xr = reinterpret_cast<float2*>(xch) + r * 4;
      xr[cg].x = mx;
      named_bar_sync(1 + q, 128);
      mx = fmaxf(fmaxf(xr[0].x, xr[1].x), fmaxf(xr[2].x, xr[3].x));
      if (!(mx > -INFINITY)) mx = 0.f;
      stamp(a, n, 5);
      // exp((s - max) / sqrt(dh)) = 2^(s c - max c), c = log2(e) / sqrt(dh): one FFMA + one EX2 per element; 2^(-inf) = 0
      const float c = scale * 1.4426950408889634f, mc = mx * c;
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 64; j++)
        if (j < cpt) {
          const float e = ex2(fmaf(__uint_as_float(p[j]), c, -mc));
          p[j] = __float_as_uint(e);
          sum += e;
        }
      xr[cg].y = sum;
      named_bar_sync(1 + q, 128);
      sum = (xr[0].y + xr[1].y) + (xr[2].y + xr[3].y);
      const float inv = (live && sum > 0.f) ? __fdividef(1.0f, sum) : 0.f;
      if (a.lse && cg == 0 && live) a.lse[(long long)bh * N + grow] = mc + log2f(sum);
      stamp(a, n, 6);
#pragma unroll
      for (int j8 = 0; j8 < 8; j8++)
        if (8 * j8 < cpt) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __uint_as_float(p[8 * j8 + i]) * inv;
          store8_split(smem + OFF_P_HI, smem + OFF_P_LO, r, col0 + 8 * j8, PCH, v);
        }
      stamp(a, n, 7);
      arrive(B_P);
      if (tid == 0) issue_o(nk, ph);
      __syncwarp();
      stamp(a, n, 8);
      // ---- the reference's atg__ / att: P to global memory, read back from the operand planes (hi + lo is the fp32 value to
      // 2^-17) while the tensor core multiplies them with V.  A warp writes 512 contiguous bytes of one row per instruction
      // (row-per-thread stores of the accumulator layout cost 32 LSU wavefronts per instruction: 11 K of 48 K cycles per tile).
      mbar_wait(&bar[B_P], ph);                      // the planes of ALL warps are complete (rows are re-dealt below)
      if (warp && a.P) {                             // warp 0 issued the MMAs meanwhile: the other 15 share the rows
        const bool vec = (N & 3) == 0;
        for (int rr = warp - 1; rr < BM; rr += TOK_WARPS - 1) {
          const int g2 = qt * BM + rr;
          if (g2 >= N) break;
          float* pg = a.P + ((long long)bh * N + g2) * N;
          for (int cc = 4 * lane; cc < N; cc += 128) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);                        // keys this tile never looked at: exact zeros
            if (cc < nk) {
              const uint32_t off = plane_off(rr, cc, PCH);
              const uint2 hw = *reinterpret_cast<const uint2*>(smem + OFF_P_HI + off), lw = *reinterpret_cast<const uint2*>(smem + OFF_P_LO + off);
              v.x = __uint_as_float(hw.x << 16) + __uint_as_float(lw.x << 16);
              v.y = __uint_as_float(hw.x & 0xffff0000u) + __uint_as_float(lw.x & 0xffff0000u);
              v.z = __uint_as_float(hw.y << 16) + __uint_as_float(lw.y << 16);
              v.w = __uint_as_float(hw.y & 0xffff0000u) + __uint_as_float(lw.y & 0xffff0000u);
            }
            if (vec) *reinterpret_cast<float4*>(pg + cc) = v;
            else {
              if (cc < N) pg[cc] = v.x;
              if (cc + 1 < N) pg[cc + 1] = v.y;
              if (cc + 2 < N) pg[cc + 2] = v.z;
              if (cc + 3 < N) pg[cc + 3] = v.w;
            }
          }
        }
      }
      stamp(a, n, 11);
      if constexpr (!PLANES) { if (it + (int)gridDim.x < items) fetch_qk(it + gridDim.x); }
      stamp(a, n, 12);
      // every warp is done reading the P planes: their tail (P lo, chunks 2 and 3) becomes the staging tile of O
      named_bar_sync(5, TOK_THREADS);
      stamp(a, n, 13);
      // ---- out = P V (+ residual)                                                               attention.py:42-46
      mbar_wait(&bar[B_O], ph);
      tc_fence_after();
      stamp(a, n, 14);
      if (PLANES && tid == 0 && it + (int)gridDim.x < items) {
        // the tensor core and every warp are done with q / k / P and v: the next item's operands land under the epilogue
        fence_async_smem();
        tma_item(it + gridDim.x);
      }
      {
        uint32_t o[16];
        tmem_ld16(lane_addr + T_O + 16 * cg, o);
#pragma unroll
        for (int k8 = 0; k8 < 2; k8++) {
          uint32_t o2[8];
          tmem_ld8(lane_addr + T_O + 64 + 16 * cg + 8 * k8, o2);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 8; j++) o[8 * k8 + j] = __float_as_uint(__uint_as_float(o[8 * k8 + j]) + __uint_as_float(o2[j]));
        }
        stamp(a, n, 9);
        // accumulator layout (row per thread) -> [128][64] fp32 tile, 16-byte units XOR-swizzled with the row
        uint8_t* ost = smem + OFF_OST;
#pragma unroll
        for (int j = 0; j < 4; j++)
          *reinterpret_cast<uint4*>(ost + r * 256 + (((4 * cg + j) ^ (r & 15)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        named_bar_sync(1 + q, 128);
        // the 128 threads of the row group write their 32 rows: 16 lanes x 16 bytes = one 256-byte row segment
        const int gt = cg * 32 + lane, unit = gt & 15;
#pragma unroll 1
        for (int ps = 0; ps < 4; ps++) {
          const int rl = 32 * q + (gt >> 4) + 8 * ps, g2 = qt * BM + rl;
          if (g2 < N) {
            const float4 ov = *reinterpret_cast<const float4*>(ost + rl * 256 + ((unit ^ (rl & 15)) << 4));
            const long long oi = ((long long)b * N + g2) * D + h * DH + 4 * unit;
            float4 rv = make_float4(0.f, 0.f, 0.f, 0.f);
            if (a.residual) rv = __ldg(reinterpret_cast<const float4*>(a.residual + oi));
            const float4 w = make_float4(ov.x + rv.x, ov.y + rv.y, ov.z + rv.z, ov.w + rv.w);
            *reinterpret_cast<float4*>(a.out + oi) = w;
            if (a.o_hi) {
              uint2 hh, ll;
              split2(w.x, w.y, hh.x, ll.x);
              split2(w.z, w.w, hh.y, ll.y);
              *reinterpret_cast<uint2*>(a.o_hi + (oi >> 1)) = hh;
              *reinterpret_cast<uint2*>(a.o_lo + (oi >> 1)) = ll;
            }
          }
        }
      }
      stamp(a, n, 10);
      // PLANES: nothing but TMA and the tensor core stands between this item and the next one's P planes, whose tail IS the
      // staging tile — a warp that is still storing its rows must not be overtaken by another row group's softmax.  (Without
      // planes every warp delivers q / k first, which orders the two.)
      if constexpr (PLANES) named_bar_sync(5, TOK_THREADS);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tbase);
  }
}

// NTB200_ATT_TCL=0 keeps the mma.sync kernels of attention_bf16.cu for these shapes.
static bool enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NTB200_ATT_TCL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on != 0;
}
}  // namespace atl


// ================================================================================================ backward
// attention.py:63-84 / gpt/attdecoderblock.py:79-99 for the same shapes.  One work item = (sample, head); the CTA walks
// the (key block kb, query tile qt) pairs of 128 x 128, key blocks from the last to the first, and per pair
//     S  = Q_t K_b^T,  dP = dO_t V_b^T                        (K-major operands, N = 128)
//     P  = 2^(S c - lse)  (recomputed: the forward's row log-sum-exp instead of the B*H*T*T matrix),  dS = P o (dP - delta) / sqrt(dh)
//     dV_b += P^T dO_t                                         (P and dO as MN-major operands: nothing is transposed)
//     dQ_t += dS K_b,  dK_b += dS^T Q_t                        (dS takes P's planes once dV has consumed them)
// with delta_i = dO_i . O_i (the softmax-backward row sum, attention.py:75).  q, k, v and dO arrive by TMA from bf16 hi / lo
// planes; all accumulators live in TMEM: S 128 + dP 128 + dK 64 + dV 64 + dQ 2 x 64 = 512 columns.
namespace atb {
using namespace um;
using atl::tma_load_3d;
using atl::tma_prefetch_3d;
using atl::ex2;

constexpr int DH = 64, BM = 128;
constexpr uint32_t PCH = BM * 128;
constexpr int TOK_WARPS = 16, THREADS = 32 * TOK_WARPS;
constexpr uint32_t OFF_Q_HI = 0, OFF_Q_LO = PCH, OFF_D_HI = 2 * PCH, OFF_D_LO = 3 * PCH, OFF_K_HI = 4 * PCH, OFF_K_LO = 5 * PCH,
                   OFF_V_HI = 6 * PCH, OFF_V_LO = 7 * PCH;
constexpr uint32_t OFF_PS_HI = 8 * PCH, OFF_PS_LO = 10 * PCH;      // P, then dS: [128 q][128 keys] = two chunks per plane
constexpr uint32_t OFF_STG = OFF_PS_HI;                            // fp32 [128][64] staging tile of the output epilogues
constexpr uint32_t OFF_DELTA = 12 * PCH, OFF_LSE = OFF_DELTA + 2 * BM * 4, OFF_BAR = OFF_LSE + 2 * BM * 4;   // [2][128] each
enum { B_TQ = 0, B_TD, B_TK, B_TV, B_SDP, B_P, B_DV, B_DS, B_DQK, B_COUNT };   // B_T*: q, dO, k, v landed (TMA)
constexpr uint32_t SMEM = OFF_BAR + B_COUNT * 8 + 16;
static_assert(SMEM <= 232448, "shared memory budget");
constexpr uint32_t T_S = 0, T_DP = 128, T_DK = 256, T_DV = 320, T_DQ = 384;

struct Args {
  const float* dout;     // [B,N,H*64]
  const float* att_out;  // [B,N,H*64]  forward output without a residual
  const float* lse;      // [B,H,N]
  float* dqkv;           // [B,N,3*H*64]
  uint32_t* g_hi;        // bf16 pairs of dqkv or null
  uint32_t* g_lo;
  int B, N, H, causal;
  long long* dbg;        // optional phase time stamps of CTA 0: [pair < 8][thread 0 / thread 480][16 events]
};
__device__ __forceinline__ void stamp(const Args& a, int n, int ev) {
  if (a.dbg && blockIdx.x == 0 && n < 8 && (threadIdx.x == 0 || threadIdx.x == 480)) a.dbg[(n * 2 + (threadIdx.x ? 1 : 0)) * 16 + ev] = clock64();
}

// D (+)= A B in three passes; operands given as descriptors of the first K = 16 step and the per-step advance in 16-byte units
__device__ __forceinline__ void mma3x(uint32_t tacc, uint64_t ah, uint64_t al, uint64_t bh, uint64_t bl, uint32_t a_step, uint32_t b_step,
                                      int ksteps, uint32_t idesc, bool fresh) {
#pragma unroll 4
  for (int ks = 0; ks < ksteps; ks++) {
    const uint64_t ao = (uint64_t)(ks * a_step), bo = (uint64_t)(ks * b_step);
    mma_bf16(tacc, al + ao, bh + bo, idesc, (fresh && ks == 0) ? 0u : 1u);
    mma_bf16(tacc, ah + ao, bl + bo, idesc, 1u);
    mma_bf16(tacc, ah + ao, bh + bo, idesc, 1u);
  }
}

__global__ void __launch_bounds__(THREADS, 1) attn_tcl_bwd_kernel(Args a, const __grid_constant__ CUtensorMap tq_hi,
                                                                  const __grid_constant__ CUtensorMap tq_lo,
                                                                  const __grid_constant__ CUtensorMap td_hi,
                                                                  const __grid_constant__ CUtensorMap td_lo) {
  extern __shared__ __align__(1024) uint8_t smem[];
  if (smem_u32(smem) & 1023u) __trap();
  float* delta_s = reinterpret_cast<float*>(smem + OFF_DELTA);
  float* lse_s = reinterpret_cast<float*>(smem + OFF_LSE);
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tslot = reinterpret_cast<uint32_t*>(bar + B_COUNT);

  pdl_launch_dependents();
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = a.N, H = a.H, D = H * DH;
  const int QT = (N + BM - 1) / BM, KT = QT, BH = a.B * H;
  if (tid == 0) {
    for (int i = 0; i < B_COUNT; i++) mbar_init(&bar[i], 1);
    mbar_init(&bar[B_P], TOK_WARPS);
    mbar_init(&bar[B_DS], TOK_WARPS);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc<512>(tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = *tslot;
  pdl_wait();

  const uint32_t sb = smem_u32(smem);
  const int q = warp & 3, cg = warp >> 2, r = 32 * q + lane;          // thread = (row of the tile = TMEM lane, column group)
  const uint32_t lane_addr = tbase + ((uint32_t)(32 * q) << 16);
  const float scale = rsqrtf((float)DH), c = scale * 1.4426950408889634f;
  auto arrive = [&](int b) {
    tc_fence_before();
    fence_async_smem();
    __syncwarp();
    if (lane == 0) mbar_arrive(&bar[b]);
  };
  // [128][64] accumulator at TMEM column tcol -> dqkv rows row0.., part 0 (dq) / 1 (dk) / 2 (dv) of head h (+ bf16 planes).
  // Accumulator layout (row per thread) -> swizzled fp32 staging tile -> 256-byte row segments, 16 lanes per row.
  auto store_tile = [&](uint32_t tcol, int part, int row0, int b, int h) {
    uint32_t o[16];
    tmem_ld16(lane_addr + tcol + 16 * cg, o);
    tmem_wait_ld();
    uint8_t* stg = smem + OFF_STG;
#pragma unroll
    for (int j = 0; j < 4; j++)
      *reinterpret_cast<uint4*>(stg + r * 256 + (((4 * cg + j) ^ (r & 15)) << 4)) = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    named_bar_sync(1 + q, 128);
    const int gt = cg * 32 + lane, unit = gt & 15;
#pragma unroll 1
    for (int ps = 0; ps < 4; ps++) {
      const int rl = 32 * q + (gt >> 4) + 8 * ps, g = row0 + rl;
      if (g < N) {
        const float4 w = *reinterpret_cast<const float4*>(stg + rl * 256 + ((unit ^ (rl & 15)) << 4));
        const long long oi = ((long long)b * N + g) * 3 * D + part * D + h * DH + 4 * unit;
        *reinterpret_cast<float4*>(a.dqkv + oi) = w;
        if (a.g_hi) {
          uint2 hh, ll;
          split2(w.x, w.y, hh.x, ll.x);
          split2(w.z, w.w, hh.y, ll.y);
          *reinterpret_cast<uint2*>(a.g_hi + (oi >> 1)) = hh;
          *reinterpret_cast<uint2*>(a.g_lo + (oi >> 1)) = ll;
        }
      }
    }
    named_bar_sync(1 + q, 128);                     // the tile is rewritten by the next store_tile / the next pair's planes
  };

  // ---- operand loads run ahead of the pairs: a tensor is re-requested as soon as the last MMA that reads the old tile has
  // retired (v after S / dP, dO after dV, q and k after dQ / dK), i.e. during the current pair, also across items.
  // Every thread keeps the same bookkeeping; thread 0 issues the copies and waits for them before the MMAs.
  enum { T_Q = 0, T_D, T_K, T_V };
  int have_item[4] = {-1, -1, -1, -1}, have_idx[4] = {-1, -1, -1, -1};
  uint32_t tph[4] = {0, 0, 0, 0};
  bool pend[4] = {false, false, false, false};
  auto request = [&](int which, int item, int idx) {
    if (have_item[which] == item && have_idx[which] == idx) return;
    have_item[which] = item; have_idx[which] = idx; pend[which] = true;
    if (tid == 0) {
      const int b = item / H, h = item % H;
      const uint32_t dst = sb + (which == T_Q ? OFF_Q_HI : which == T_D ? OFF_D_HI : which == T_K ? OFF_K_HI : OFF_V_HI);
      const int col = (which == T_Q ? 0 : which == T_D ? 0 : which == T_K ? D : 2 * D) + h * DH;
      fence_async_smem();
      mbar_expect_tx(&bar[B_TQ + which], 2u * PCH);
      tma_load_3d(dst, which == T_D ? &td_hi : &tq_hi, &bar[B_TQ + which], col, idx * BM, b);
      tma_load_3d(dst + PCH, which == T_D ? &td_lo : &tq_lo, &bar[B_TQ + which], col, idx * BM, b);
    }
  };
  auto await = [&](int which) {                      // thread 0 blocks; everybody flips the phase
    if (!pend[which]) return;
    if (tid == 0) mbar_wait(&bar[B_TQ + which], tph[which]);
    tph[which] ^= 1;
    pend[which] = false;
  };
  auto advance = [&](int& item, int& kb, int& qt) {   // next (item, key block, query tile); item >= BH: none
    qt--;
    if (qt < (a.causal ? kb : 0)) { kb--; qt = QT - 1; }
    if (kb < 0) { item += gridDim.x; kb = KT - 1; qt = QT - 1; }
  };

  // delta_i = dO_i . O_i (attention.py:75) and the row log-sum-exp of query tile (item, qt) -> buffer `buf`:
  // 4 lanes per row, 256 contiguous bytes
  auto compute_delta = [&](int item_, int qt_, int buf) {
    const int row = tid >> 2, qq = tid & 3, g = qt_ * BM + row;
    const int b_ = item_ / H, h_ = item_ % H;
    float acc = 0.f;
    if (g < N) {
      const long long oi = ((long long)b_ * N + g) * D + h_ * DH + 16 * qq;
      const float4* pd = reinterpret_cast<const float4*>(a.dout + oi);
      const float4* po = reinterpret_cast<const float4*>(a.att_out + oi);
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const float4 x = __ldg(pd + j), y = __ldg(po + j);
        acc += x.x * y.x + x.y * y.y + x.z * y.z + x.w * y.w;
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (qq == 0) delta_s[buf * BM + row] = acc;
    if (tid < BM) lse_s[buf * BM + tid] = (qt_ * BM + tid < N) ? __ldg(a.lse + (long long)item_ * N + qt_ * BM + tid) : 0.f;
  };
  int dbuf = 0;
  bool dnext = false;
  uint32_t pp = 0;                                  // phase of the per-pair barriers
  int np = 0;                                       // pairs done by this CTA (time stamps)
  int item = blockIdx.x, kb = KT - 1, qt = QT - 1;
  int d_item = -1, d_qt = -1;                       // query tile whose delta / lse sit in shared memory
  if (item < BH) { request(T_Q, item, qt); request(T_K, item, kb); request(T_D, item, qt); request(T_V, item, kb); }
  while (item < BH) {
    const int b = item / H, h = item % H;
    int n_item = item, n_kb = kb, n_qt = qt;
    advance(n_item, n_kb, n_qt);
    const bool has_next = n_item < BH;
    const bool first_k = (qt == QT - 1);                         // first pair of this key block: dK, dV start from zero
    const bool first_q = (kb == (a.causal ? qt : KT - 1));       // first pair of this query tile: dQ starts from zero
    const bool last_k = (qt == (a.causal ? kb : 0));             // last pair of this key block: dK, dV are complete
    const bool last_item = last_k && kb == 0;
    stamp(a, np, 0);
    if (tid == 0 && has_next) {
      // the tiles the next pair will ask for, towards L2 now (a [128][64] tile of a plane is 128 rows of 128 bytes, 3D * 2
      // bytes apart: 4.4 K cycles when it comes from DRAM at the moment it is needed)
      const int nb = n_item / H, nc = (n_item % H) * DH;
      if (n_item != item || n_qt != qt) {
        tma_prefetch_3d(&tq_hi, nc, n_qt * BM, nb); tma_prefetch_3d(&tq_lo, nc, n_qt * BM, nb);
        tma_prefetch_3d(&td_hi, nc, n_qt * BM, nb); tma_prefetch_3d(&td_lo, nc, n_qt * BM, nb);
      }
      if (n_item != item || n_kb != kb) {
        tma_prefetch_3d(&tq_hi, D + nc, n_kb * BM, nb); tma_prefetch_3d(&tq_lo, D + nc, n_kb * BM, nb);
        tma_prefetch_3d(&tq_hi, 2 * D + nc, n_kb * BM, nb); tma_prefetch_3d(&tq_lo, 2 * D + nc, n_kb * BM, nb);
      }
    }
    // ---- S = Q K^T, dP = dO V^T
    {
      const uint32_t id = idesc_bf16(128, 128, 0, 0);
      await(T_Q);
      await(T_K);
      if (tid == 0) {
        tc_fence_after();
        mma3x(tbase + T_S, desc_kmajor(sb + OFF_Q_HI), desc_kmajor(sb + OFF_Q_LO), desc_kmajor(sb + OFF_K_HI), desc_kmajor(sb + OFF_K_LO),
              KMAJ_STEP >> 4, KMAJ_STEP >> 4, 4, id, true);
      }
      await(T_D);
      await(T_V);
      if (tid == 0) {
        tc_fence_after();
        mma3x(tbase + T_DP, desc_kmajor(sb + OFF_D_HI), desc_kmajor(sb + OFF_D_LO), desc_kmajor(sb + OFF_V_HI), desc_kmajor(sb + OFF_V_LO),
              KMAJ_STEP >> 4, KMAJ_STEP >> 4, 4, id, true);
        mma_commit(&bar[B_SDP]);
      }
      __syncwarp();
    }
    stamp(a, np, 1);
    // ---- delta / lse of this query tile: normally computed one pair ahead (below); only the first pair computes here
    if (d_item != item || d_qt != qt) {
      dbuf ^= 1;
      compute_delta(item, qt, dbuf);
      d_item = item; d_qt = qt;
      named_bar_sync(5, THREADS);
    }
    stamp(a, np, 2);
    // ---- P and dS of this thread's 32 key columns
    uint32_t pv[32], dv[32];
    mbar_wait(&bar[B_SDP], pp);
    tc_fence_after();
    if (has_next) request(T_V, n_item, n_kb);        // dP has retired: v may go
    tmem_ld32(lane_addr + T_S + 32 * cg, pv);
    tmem_ld32(lane_addr + T_DP + 32 * cg, dv);
    tmem_wait_ld();
    stamp(a, np, 3);
    {
      const int g = qt * BM + r, cb = kb * BM + 32 * cg;
      const bool live = g < N;
      const float ls = lse_s[dbuf * BM + r], dl = delta_s[dbuf * BM + r];
      const bool edge = (cb + 31 >= N) || (a.causal && cb + 31 > qt * BM + 32 * q);
#pragma unroll
      for (int j = 0; j < 32; j++) {
        bool ok = live;
        if (edge) ok = ok && (cb + j < N) && !(a.causal && cb + j > g);
        const float p = ok ? ex2(fmaf(__uint_as_float(pv[j]), c, -ls)) : 0.f;
        pv[j] = __float_as_uint(p);
        dv[j] = __float_as_uint(p * (__uint_as_float(dv[j]) - dl) * scale);
      }
    }
#pragma unroll
    for (int j8 = 0; j8 < 4; j8++) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = __uint_as_float(pv[8 * j8 + i]);
      store8_split(smem + OFF_PS_HI, smem + OFF_PS_LO, r, 32 * cg + 8 * j8, PCH, v);
    }
    stamp(a, np, 4);
    arrive(B_P);
    // ---- dV += P^T dO
    if (tid == 0) {
      mbar_wait(&bar[B_P], pp);
      tc_fence_after();
      mma3x(tbase + T_DV, desc_mnmajor(sb + OFF_PS_HI, PCH), desc_mnmajor(sb + OFF_PS_LO, PCH), desc_mnmajor(sb + OFF_D_HI, 8192),
            desc_mnmajor(sb + OFF_D_LO, 8192), MNMAJ_STEP >> 4, MNMAJ_STEP >> 4, 8, idesc_bf16(128, 64, 1, 1), first_k);
      mma_commit(&bar[B_DV]);
    }
    __syncwarp();
    mbar_wait(&bar[B_DV], pp);                       // the tensor core has consumed P: its planes take dS
    tc_fence_after();
    if (has_next) request(T_D, n_item, n_qt);        // dP and dV have retired: dO may go
    stamp(a, np, 5);
#pragma unroll
    for (int j8 = 0; j8 < 4; j8++) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; i++) v[i] = __uint_as_float(dv[8 * j8 + i]);
      store8_split(smem + OFF_PS_HI, smem + OFF_PS_LO, r, 32 * cg + 8 * j8, PCH, v);
    }
    stamp(a, np, 6);
    arrive(B_DS);
    // ---- dQ += dS K, dK += dS^T Q
    if (tid == 0) {
      mbar_wait(&bar[B_DS], pp);
      tc_fence_after();
      // dS as a K-major A operand: K = 16 step ks lives in chunk ks / 4 at 32 (ks % 4) bytes
      const uint64_t ah0 = desc_kmajor(sb + OFF_PS_HI), al0 = desc_kmajor(sb + OFF_PS_LO);
      const uint64_t bh0 = desc_mnmajor(sb + OFF_K_HI, 8192), bl0 = desc_mnmajor(sb + OFF_K_LO, 8192);
      const uint32_t idq = idesc_bf16(128, 64, 0, 1);
#pragma unroll 4
      for (int ks = 0; ks < 8; ks++) {
        const uint64_t pa = (uint64_t)((ks >> 2) * (PCH >> 4) + (ks & 3) * (KMAJ_STEP >> 4)), pb = (uint64_t)(ks * (MNMAJ_STEP >> 4));
        mma_bf16(tbase + T_DQ + 64 * qt, al0 + pa, bh0 + pb, idq, (first_q && ks == 0) ? 0u : 1u);
        mma_bf16(tbase + T_DQ + 64 * qt, ah0 + pa, bl0 + pb, idq, 1u);
        mma_bf16(tbase + T_DQ + 64 * qt, ah0 + pa, bh0 + pb, idq, 1u);
      }
      mma3x(tbase + T_DK, desc_mnmajor(sb + OFF_PS_HI, PCH), desc_mnmajor(sb + OFF_PS_LO, PCH), desc_mnmajor(sb + OFF_Q_HI, 8192),
            desc_mnmajor(sb + OFF_Q_LO, 8192), MNMAJ_STEP >> 4, MNMAJ_STEP >> 4, 8, idesc_bf16(128, 64, 1, 1), first_k);
      mma_commit(&bar[B_DQK]);
    }
    __syncwarp();
    if (has_next && (n_item != item || n_qt != qt)) {
      // the next pair works on another query tile: its delta / lse into the other buffer while dQ and dK are being multiplied
      // (everybody is past its reads of the buffers: B_DS was complete before the MMAs were issued ... for the OTHER buffer the
      // last readers finished a whole pair ago)
      compute_delta(n_item, n_qt, dbuf ^ 1);
      d_item = n_item; d_qt = n_qt;
      dnext = true;
    }
    mbar_wait(&bar[B_DQK], pp);
    tc_fence_after();
    if (has_next) { request(T_K, n_item, n_kb); request(T_Q, n_item, n_qt); }     // everything has retired
    stamp(a, np, 7);
    pp ^= 1;
    if (last_k) {                                    // dK, dV of this key block are complete
      store_tile(T_DK, 1, kb * BM, b, h);
      store_tile(T_DV, 2, kb * BM, b, h);
      stamp(a, np, 8);
    }
    if (last_item) {
      for (int t = 0; t < QT; t++) store_tile(T_DQ + 64 * t, 0, t * BM, b, h);
      stamp(a, np, 9);
    }
    if (last_k || dnext) named_bar_sync(5, THREADS); // the staging tile is the P / dS planes of the next pair; delta is published
    if (dnext) { dbuf ^= 1; dnext = false; }
    np++;
    item = n_item; kb = n_kb; qt = n_qt;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<512>(tbase);
  }
}
}  // namespace atb

bool attention_tcl_ok(int N, int H, int dh, int mask_kind, bool want_scores) {
  return atl::enabled() && dh == 64 && N > 64 && N <= atl::MAXK && H >= 1 && mask_kind != NT_MASK_DENSE && !want_scores &&
         g_gemm_mode >= 1;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// bf16 plane of qkv [B][N][3D] as a 3-D tensor {3D, N, B}; box {64 columns, 128 rows, 1}: rows past N are zero-filled
static int plane_map(CUtensorMap* tm, const void* base, int B, int N, int D3) {
  static EncodeTiledFn enc = nullptr;
  if (!enc) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    NT_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    NT_REQUIRE(qres == cudaDriverEntryPointSuccess && fn, "cuTensorMapEncodeTiled is not available from the driver");
    enc = (EncodeTiledFn)fn;
  }
  cuuint64_t dims[3] = {(cuuint64_t)D3, (cuuint64_t)N, (cuuint64_t)B};
  cuuint64_t strides[2] = {(cuuint64_t)D3 * 2, (cuuint64_t)N * D3 * 2};
  cuuint32_t box[3] = {64, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(qkv plane) failed (%d) B=%d N=%d 3D=%d", (int)r, B, N, D3);
  return 0;
}

int attention_fwd_tcl(const float* qkv, const void* qkv_hi, const void* qkv_lo, int mask_kind, float* out, float* P, int B, int N,
                      int H, const float* residual, void* o_hi, void* o_lo, float* lse) {
  static bool cfg = false;
  if (!cfg) {
    NT_CHECK(cudaFuncSetAttribute(atl::attn_tcl_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atl::SMEM));
    NT_CHECK(cudaFuncSetAttribute(atl::attn_tcl_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atl::SMEM));
    cfg = true;
  }
  atl::Args a{};
  a.qkv = qkv; a.out = out; a.P = P; a.residual = residual; a.o_hi = (uint32_t*)o_hi; a.o_lo = (uint32_t*)o_lo;
  a.B = B; a.N = N; a.H = H; a.causal = mask_kind == NT_MASK_CAUSAL ? 1 : 0;
  a.dbg = atl::g_dbg;
  a.lse = lse;
  CUtensorMap tm_hi, tm_lo;
  memset(&tm_hi, 0, sizeof(tm_hi));
  memset(&tm_lo, 0, sizeof(tm_lo));
  a.planes = (qkv_hi && qkv_lo && !((uintptr_t)qkv_hi & 15) && !((uintptr_t)qkv_lo & 15)) ? 1 : 0;
  if (a.planes && (plane_map(&tm_hi, qkv_hi, B, N, 3 * H * atl::DH) || plane_map(&tm_lo, qkv_lo, B, N, 3 * H * atl::DH))) return 1;
  const int items = B * H * ((N + atl::BM - 1) / atl::BM);
  const int grid = items < g_sm_count ? items : g_sm_count;
  if (a.planes) NT_CHECK(launch_k(atl::attn_tcl_fwd_kernel<true>, dim3(grid), dim3(atl::THREADS), atl::SMEM, a, tm_hi, tm_lo));
  else NT_CHECK(launch_k(atl::attn_tcl_fwd_kernel<false>, dim3(grid), dim3(atl::THREADS), atl::SMEM, a, tm_hi, tm_lo));
  NT_LAUNCH_OK();
  return 0;
}


// NTB200_ATT_TCL_BWD=0 keeps the mma.sync backward.
bool attention_bwd_tcl_ok(int N, int H, int dh, int mask_kind) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("NTB200_ATT_TCL_BWD"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && atl::enabled() && dh == 64 && N > 64 && N <= atl::MAXK && H >= 1 && mask_kind != NT_MASK_DENSE && g_gemm_mode >= 1;
}

int attention_bwd_tcl(const float* dout, const void* dout_hi, const void* dout_lo, const void* qkv_hi, const void* qkv_lo,
                      const float* att_out, const float* lse, float* dqkv, void* g_hi, void* g_lo, int B, int N, int H, int causal) {
  static bool cfg = false;
  if (!cfg) {
    NT_CHECK(cudaFuncSetAttribute(atb::attn_tcl_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)atb::SMEM));
    cfg = true;
  }
  NT_REQUIRE(!((uintptr_t)dout_hi & 15) && !((uintptr_t)dout_lo & 15) && !((uintptr_t)qkv_hi & 15) && !((uintptr_t)qkv_lo & 15),
             "attention_bwd_tcl: planes must be 16-byte aligned");
  atb::Args a{};
  a.dout = dout; a.att_out = att_out; a.lse = lse; a.dqkv = dqkv; a.g_hi = (uint32_t*)g_hi; a.g_lo = (uint32_t*)g_lo;
  a.B = B; a.N = N; a.H = H; a.causal = causal;
  a.dbg = atl::g_dbg_bwd;
  CUtensorMap tq_hi, tq_lo, td_hi, td_lo;
  if (plane_map(&tq_hi, qkv_hi, B, N, 3 * H * atb::DH) || plane_map(&tq_lo, qkv_lo, B, N, 3 * H * atb::DH) ||
      plane_map(&td_hi, dout_hi, B, N, H * atb::DH) || plane_map(&td_lo, dout_lo, B, N, H * atb::DH))
    return 1;
  const int items = B * H;
  const int grid = items < g_sm_count ? items : g_sm_count;
  NT_CHECK(launch_k(atb::attn_tcl_bwd_kernel, dim3(grid), dim3(atb::THREADS), atb::SMEM, a, tq_hi, tq_lo, td_hi, td_lo));
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt

/* development aid: phase time stamps of CTA 0 ([8 items][2 threads][16 events] clock64 values); NULL switches them off */
extern "C" int nt_attention_tcl_debug(long long* stamps) {
  nt::atl::g_dbg = stamps;
  return 0;
}
extern "C" int nt_attention_tcl_bwd_debug(long long* stamps) {
  nt::atl::g_dbg_bwd = stamps;
  return 0;
}
