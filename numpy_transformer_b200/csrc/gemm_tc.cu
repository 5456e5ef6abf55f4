// tcgen05 / TMEM / TMA GEMM for sm_100a:  C[M,N] = alpha * op(A) op(B)  (+ shared epilogue)
//
//   * operands are the fp32 tensors as they sit in HBM; TMA (cp.async.bulk.tensor) lands 128-byte
//     swizzled tiles in shared memory in exactly the canonical UMMA layout, tcgen05.mma kind::tf32
//     consumes them straight from smem, the fp32 accumulator lives in TMEM.
//   * K-major operands (contraction dim contiguous: X in X@W, dY and W in dY@W^T) use SWIZZLE_128B;
//     MN-major operands (W in X@W, X and dY in X^T@dY) use SWIZZLE_128B_BASE32B, the only MN-major
//     layout UMMA accepts for 32-bit elements.  No operand is ever transposed or converted in HBM.
//   * passes == 3 (gemm mode 1): error-compensated 3xTF32.  Four converter warps split every landed
//     tile into hi = tf32(x) and lo = x - hi in shared memory (element-wise, so layout-agnostic) and
//     the issuer accumulates hi*hi + hi*lo + lo*hi  ->  ~2^-21 relative error, fp32-class results
//     (rel <= 1e-4 bar of the spec; single-pass TF32 measures ~3e-3 on gradients and fails it).
//   * warp roles: w0 = TMA producer, w1 = TMEM allocator + single-thread MMA issuer,
//     w2..w5 = converters during the main loop, then the epilogue (tcgen05.ld -> bias / GELU / ReLU /
//     act' / residual -> global, or red.add for the split-K weight-gradient GEMM).
//   * one 128 x BN output tile per CTA, BN in {32,64,96,128}; blockIdx.z splits K (dW: K = rows of
//     the batch, thousands long) so the grid fills the 148 SMs; ragged M/N/K edges come from TMA
//     zero fill + guarded stores.
#include "common.cuh"
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdlib>

namespace nt {

static constexpr int BM = 128;      // UMMA M (cta_group::1)
static constexpr int BK = 32;       // fp32 elements per k-block = one 128-byte swizzle row
static constexpr int TC_WORKERS = 8;                    // converter / epilogue warps
static constexpr int TC_THREADS = 64 + 32 * TC_WORKERS;

struct TcParams {
  int M, N, K;            // problem (C is M x N, contraction K)
  int BN;                 // N tile
  int a_mn, b_mn;         // 1 = MN-major operand
  int passes;             // 1 or 3
  int stages;
  int kblocks_per_split;  // k-blocks handled by one blockIdx.z
  int dbg_skip_a;         // experiment: do not load A (results are garbage)
  int dbg_mma;            // experiment: MMAs issued per k-block (default 4)
  int a_3d, b_3d;         // MN-major operand described by a 3-D map {32, k, chunk}: one TMA per k-block
  int bf16;               // 1: operands are pre-split bf16 hi/lo K-major copies (4 tensor maps), kind::f16, 64 k per stage
  int wide;               // bf16 only: two MMAs per K = 16 step instead of three — A.hi x [B.hi ; B.lo] as ONE instruction of
                          // N = 2 BN (the lo tile sits right behind the hi tile, so the pair is one operand) and A.lo x B.hi;
                          // the accumulator has 2 BN columns, summed by the epilogue
  long long ldc;
  float* C;
  Epilogue ep;
};

// ---------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  // bounded spin: a protocol bug traps (process error) instead of hanging the GPU
#pragma unroll 1
  for (uint32_t it = 0; it < (1u << 24); ++it) {
    asm volatile(
        "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}"
      ::"r"(tmem_c), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
      : "memory");
}
// bf16 hi/lo split of two fp32 values (same arithmetic as split_bf16.cu)
__device__ __forceinline__ void tc_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// 64-bit shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
//   [0,14) start>>4 | [16,30) LBO>>4 | [32,46) SBO>>4 | [46,48) version=1 | [61,64) layout type
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}
static constexpr uint32_t LAYOUT_SW128 = 2;          // K-major tiles
static constexpr uint32_t LAYOUT_SW128_BASE32B = 1;  // MN-major 32-bit tiles

#ifdef NT_TC_TIMING
__device__ unsigned long long g_tc_times[16384 * 8];
#define TC_STAMP(ev)                                                                         \
  do {                                                                                       \
    unsigned long long _t;                                                                   \
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(_t));                                   \
    int _cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);               \
    if (_cta < 16384) g_tc_times[_cta * 8 + (ev)] = _t;                                       \
  } while (0)
#else
#define TC_STAMP(ev) do {} while (0)
#endif

// ---------------------------------------------------------------- kernel
// Epilogue flags are template parameters: a runtime-configured epilogue costs ~8 uniform branches per
// 4-row store group, which measured ~250 cycles per group on a path that runs once per CTA.
template <int ACT, int DACT, bool PRE_OUT, bool ATOMIC, bool ADDEND>
// The GELU-derivative epilogue would take 68-71 registers: with 320 threads that is one CTA too many for the three co-resident
// single-stage CTAs the forward / dX configuration runs with (65536 / (3 x 320) = 68), so those instances are capped at 64.
__global__ void __launch_bounds__(TC_THREADS, (DACT == NT_ACT_GELU && !ATOMIC) ? 3 : 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmAl, const __grid_constant__ CUtensorMap tmBl, const TcParams p) {
  // the two-MMA product (TcParams::wide) is only compiled into the weight-gradient instance: it needs 2 BN TMEM columns,
  // which would cut the forward / dX configuration from three co-resident CTAs to two (measured: dW -18 %, forward +15 %),
  // and its epilogue code pushes the other instances over the 68 registers three CTAs per SM allow
  const bool wide = ATOMIC && p.wide;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages] x { A hi | B hi | (A lo | B lo) }, then barriers
  // offset arithmetic on the __shared__ array (not on a generic uintptr_t) keeps the address space known,
  // so the converter / epilogue staging compile to LDS/STS instead of generic LD/ST
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t a_bytes = BM * BK * 4;            // 16 KB
  const uint32_t b_bytes = p.BN * BK * 4;
  const uint32_t stage_bytes = (a_bytes + b_bytes) * (p.passes == 3 ? 2 : 1);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + p.stages;
  uint64_t* conv = bars + 2 * p.stages;
  uint64_t* tmem_full = bars + 3 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * p.stages + 1);
  float* cs_s = reinterpret_cast<float*>(tmem_slot + 4);      // [128] column sums of the B tile (db fusion)

  pdl_launch_dependents();            // the next kernel may start its own prologue now
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) TC_STAMP(0);
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * p.BN;
  const int bke = p.bf16 ? 2 * BK : BK;        // elements per 128-byte k-block row (bf16: 64, fp32: 32)
  const int total_kb = (p.K + bke - 1) / bke;
  const int kb_begin = blockIdx.z * p.kblocks_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kblocks_per_split);
  const int nkb = kb_end - kb_begin;       // >= 1 by construction of the grid
  uint32_t tmem_cols = 32;
  while ((int)tmem_cols < (wide ? 2 * p.BN : p.BN)) tmem_cols <<= 1;

  const uint32_t a_bytes_ = BM * BK * 4;
  // one stage's TMA traffic (lambda so the pre-sync prologue and the steady-state loop share it)
  auto issue_stage = [&](int i) {
    const int s = i % p.stages;
    uint8_t* st = smem + (size_t)s * stage_bytes;
    const uint32_t sa = smem_u32(st), sb = sa + a_bytes_;
    const int k0 = (kb_begin + i) * bke;
    if (p.bf16) {
      // pre-split operands: hi and lo tiles of A and B arrive by TMA (no in-kernel conversion)
      mbar_expect_tx(&full[s], 2 * (a_bytes_ + b_bytes));
      const uint32_t sbl_ = sb + b_bytes, sal_ = sbl_ + b_bytes;          // A.hi | B.hi | B.lo | A.lo
      if (!p.a_mn) {
        tma_load_2d(sa, &tmA, &full[s], k0, m0);                            // box {64 k, 128 m} bf16
        tma_load_2d(sal_, &tmAl, &full[s], k0, m0);
      } else {
        // MN-major: [k rows][64 m] boxes of 64 x 128 B, one per 64-wide chunk of the tile (chunk stride 8 KB);
        // a 3-D map {64, k, chunk} lands all chunks with one instruction when the extent is a multiple of 64
        if (p.a_3d) {
          tma_load_3d(sa, &tmA, &full[s], 0, k0, m0 >> 6);
          tma_load_3d(sal_, &tmAl, &full[s], 0, k0, m0 >> 6);
        } else {
          for (int j = 0; j < BM / 64; j++) {
            tma_load_2d(sa + j * 8192, &tmA, &full[s], m0 + 64 * j, k0);
            tma_load_2d(sal_ + j * 8192, &tmAl, &full[s], m0 + 64 * j, k0);
          }
        }
      }
      if (!p.b_mn) {
        tma_load_2d(sb, &tmB, &full[s], k0, n0);                            // box {64 k, BN n}
        tma_load_2d(sbl_, &tmBl, &full[s], k0, n0);
      } else {
        if (p.b_3d) {
          tma_load_3d(sb, &tmB, &full[s], 0, k0, n0 >> 6);
          tma_load_3d(sbl_, &tmBl, &full[s], 0, k0, n0 >> 6);
        } else {
          for (int j = 0; j < p.BN / 64; j++) {
            tma_load_2d(sb + j * 8192, &tmB, &full[s], n0 + 64 * j, k0);
            tma_load_2d(sbl_ + j * 8192, &tmBl, &full[s], n0 + 64 * j, k0);
          }
        }
      }
      return;
    }
    mbar_expect_tx(&full[s], (p.dbg_skip_a ? 0 : a_bytes_) + b_bytes);
    if (p.dbg_skip_a) {
    } else if (!p.a_mn) {
      tma_load_2d(sa, &tmA, &full[s], k0, m0);                       // box {32 k, 128 m}
    } else if (p.a_3d) {
      tma_load_3d(sa, &tmA, &full[s], 0, k0, m0 >> 5);               // box {32 m, 32 k, 4 chunks} in one instruction
    } else {
      for (int j = 0; j < BM / 32; j++) tma_load_2d(sa + j * (BK * 128), &tmA, &full[s], m0 + 32 * j, k0);  // box {32 m, 32 k}
    }
    if (!p.b_mn) {
      tma_load_2d(sb, &tmB, &full[s], k0, n0);                       // box {32 k, BN n}
    } else if (p.b_3d) {
      tma_load_3d(sb, &tmB, &full[s], 0, k0, n0 >> 5);
    } else {
      for (int j = 0; j < p.BN / 32; j++) tma_load_2d(sb + j * (BK * 128), &tmB, &full[s], n0 + 32 * j, k0);
    }
  };
  // db = colsum(dY) rides along with dW = x^T dY: the worker warps already sweep every landed B tile
  const bool do_colsum = (p.ep.colsum != nullptr) && p.b_mn && blockIdx.y == 0;
  if (threadIdx.x < 128) cs_s[threadIdx.x] = 0.f;
  const int prologue = nkb < p.stages ? nkb : p.stages;
  if (threadIdx.x == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
    if (p.bf16) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmAl)) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmBl)) : "memory");
    }
    for (int s = 0; s < p.stages; s++) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], (p.passes == 1 && do_colsum) ? 1 + TC_WORKERS : 1);
      mbar_init(&conv[s], TC_WORKERS);
    }
    mbar_init(tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    pdl_wait();                        // operands are outputs of the preceding kernels
    // first ring of loads goes out before the CTA-wide sync: the DRAM/L2 latency of the first tiles
    // overlaps the TMEM allocation and barrier hand-shake (all slots are trivially empty)
    for (int i = 0; i < prologue; i++) issue_stage(i);
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  pdl_wait();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) TC_STAMP(1);

  if (warp == 0) {
    // ===================== TMA producer (one lane) =====================
    if (lane == 0) {
      for (int i = prologue; i < nkb; i++) {
        const int s = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait(&empty[s], ph ^ 1);
        issue_stage(i);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (one lane) =====================
    if (lane == 0) {
      // instruction descriptor: D=f32, A=B=tf32, majors, N>>3, M>>4
      // (A/B format: 2 = TF32 for kind::tf32; 1 = BF16 for kind::f16)
      const uint32_t fmt = p.bf16 ? 1u : 2u;
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      // per-MMA (K = 8 fp32) start-address advance and descriptor strides
      // MN-major canonical layouts (cute mma_traits_sm100.hpp): 64-wide MN chunks of [k rows][128 B]; LBO = chunk
      // stride, SBO = stride between swizzle atoms along K (8 k-rows for 16-bit SW128, 4 for the 32-bit BASE32B atom)
      const uint32_t mn_adv = p.bf16 ? 16 * 128 : 8 * 128, mn_lbo = p.bf16 ? 64 * 128 : BK * 128, mn_sbo = p.bf16 ? 1024 : 512;
      const uint32_t mn_lay = p.bf16 ? LAYOUT_SW128 : LAYOUT_SW128_BASE32B;
      const uint32_t a_adv = p.a_mn ? mn_adv : 32, b_adv = p.b_mn ? mn_adv : 32;
      const uint32_t a_lbo = p.a_mn ? mn_lbo : 16, a_sbo = p.a_mn ? mn_sbo : 1024;
      const uint32_t b_lbo = p.b_mn ? mn_lbo : 16, b_sbo = p.b_mn ? mn_sbo : 1024;
      const uint32_t a_lay = p.a_mn ? mn_lay : LAYOUT_SW128;
      const uint32_t b_lay = p.b_mn ? mn_lay : LAYOUT_SW128;
      const uint32_t idesc_w = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                               ((uint32_t)((2 * p.BN) >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
      uint32_t accum = 0;
      for (int i = 0; i < nkb; i++) {
        const int s = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait((p.passes == 3 && !p.bf16) ? &conv[s] : &full[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes), sb = sa + a_bytes;
        // lo tiles: fp32 3-pass keeps A.lo | B.lo behind the hi pair; bf16 puts B.lo right behind B.hi (see TcParams::wide)
        const uint32_t sbl = p.bf16 ? sb + b_bytes : sb + b_bytes + a_bytes;
        const uint32_t sal = p.bf16 ? sbl + b_bytes : sb + b_bytes;
#pragma unroll
        for (int k = 0; k < BK / 8; k++) {
          if (k >= p.dbg_mma && p.dbg_mma < 4) break;
          const uint64_t ah = make_smem_desc(sa + k * a_adv, a_lbo, a_sbo, a_lay);
          const uint64_t bh = make_smem_desc(sb + k * b_adv, b_lbo, b_sbo, b_lay);
          if (p.bf16) {
            // 16 bf16 per MMA = the same 32-byte advance as 8 fp32: descriptors are identical to the K-major fp32 case
            const uint64_t al = make_smem_desc(sal + k * a_adv, a_lbo, a_sbo, a_lay);
            const uint64_t bl = make_smem_desc(sbl + k * b_adv, b_lbo, b_sbo, b_lay);
            if (wide) {
              // two instructions of N = 2 BN and BN instead of three of BN (tools/umma_rate.cu: N = 128 66, N = 256 124 cycles)
              umma_f16(tmem_base, ah, bh, idesc_w, accum);   // [hi*hi | hi*lo] -> columns [0, BN) and [BN, 2 BN)
              umma_f16(tmem_base, al, bh, idesc, 1);         // lo*hi -> columns [0, BN)
            } else {
              umma_f16(tmem_base, al, bh, idesc, accum);       // small terms first
              umma_f16(tmem_base, ah, bl, idesc, 1);
              umma_f16(tmem_base, ah, bh, idesc, 1);
            }
          } else if (p.passes == 3) {
            const uint64_t al = make_smem_desc(sal + k * a_adv, a_lbo, a_sbo, a_lay);
            const uint64_t bl = make_smem_desc(sbl + k * b_adv, b_lbo, b_sbo, b_lay);
            umma_tf32(tmem_base, al, bh, idesc, accum);      // small terms first
            umma_tf32(tmem_base, ah, bl, idesc, 1);
            umma_tf32(tmem_base, ah, bh, idesc, 1);
          } else {
            umma_tf32(tmem_base, ah, bh, idesc, accum);
          }
          accum = 1;
        }
        umma_commit(&empty[s]);          // frees the smem slot when these MMAs retire
      }
      umma_commit(tmem_full);            // accumulator complete
      TC_STAMP(5);
    }
    __syncwarp();
  } else {
    // ===================== converters (3xTF32 split), then epilogue =====================
    const int ct = threadIdx.x - 64;     // 0 .. 32*TC_WORKERS-1
    float4 cs[4];
#pragma unroll
    for (int j = 0; j < 4; j++) cs[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    if ((p.passes == 3 && !p.bf16) || do_colsum) {
      const uint32_t hi_bytes = a_bytes + b_bytes;
      for (int i = 0; i < nkb; i++) {
        const int s = i % p.stages;
        const uint32_t ph = (i / p.stages) & 1;
        mbar_wait(&full[s], ph);
        if (ct == 0 && i == 0) TC_STAMP(2);
        const float4* hi = reinterpret_cast<const float4*>(smem + (size_t)s * stage_bytes);
        if (p.passes == 3) {
          float4* lo = reinterpret_cast<float4*>(smem + (size_t)s * stage_bytes + hi_bytes);
          const int n4 = hi_bytes / 16;
#pragma unroll 4
          for (int e = ct; e < n4; e += 32 * TC_WORKERS) {
            // kind::tf32 ignores the low 13 mantissa bits of its operands (measured: matches a truncated
            // reference to 2.6e-7, a rounded one only to 1e-3), so the landed tile IS hi; only lo is written.
            const float4 x = hi[e];
            float4 l;
            l.x = x.x - __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u);
            l.y = x.y - __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u);
            l.z = x.z - __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u);
            l.w = x.w - __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u);
            lo[e] = l;
          }
        }
        if (do_colsum) {
          // thread ct owns (k-row ct/8, 16-byte chunk ct%8) of every 32-column box of the B tile
          const float4* bt = hi + a_bytes / 16;
#pragma unroll
          for (int j = 0; j < 4; j++)
            if (j * 32 < p.BN) {
              const float4 v = bt[j * 256 + ct];
              cs[j].x += v.x; cs[j].y += v.y; cs[j].z += v.z; cs[j].w += v.w;
            }
        }
        if (p.passes == 3) {
          if (p.dbg_mma != 5) fence_async_smem();   // generic-proxy writes -> visible to the tensor core (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(&conv[s]);
        } else {
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[s]);    // single pass: the stage is free once MMA *and* the sweep are done
        }
        if (ct == 0 && i == 0) TC_STAMP(3);
      }
    }
    if (do_colsum) {
      // Combine the per-thread partials.  The same lane of the 8 worker warps holds the same columns (k-rows
      // q, q+4, ..), so first sum across warps through shared memory (the pipeline stages are idle once the
      // accumulator is complete), then un-swizzle (Swizzle<2,5,2>: 32-byte chunk ^= k-row & 3) with a 4-way
      // shared atomic, then one red.add per column per CTA.
      mbar_wait(tmem_full, 0);
      float4* red = reinterpret_cast<float4*>(smem);          // [box][warp][lane]
      const int wq = warp - 2;
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (j * 32 < p.BN) red[(j * TC_WORKERS + wq) * 32 + lane] = cs[j];
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_WORKERS) : "memory");
      if (ct < p.BN) {
        const int j = ct >> 5, ln = ct & 31;
        float4 tot = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int w = 0; w < TC_WORKERS; w++) {
          const float4 v = red[(j * TC_WORKERS + w) * 32 + ln];
          tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        const int q = ln >> 3, c16 = ln & 7;
        const int colb = j * 32 + (((c16 >> 1) ^ q) << 3) + ((c16 & 1) << 2);
        atomicAdd(&cs_s[colb + 0], tot.x);
        atomicAdd(&cs_s[colb + 1], tot.y);
        atomicAdd(&cs_s[colb + 2], tot.z);
        atomicAdd(&cs_s[colb + 3], tot.w);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(32 * TC_WORKERS) : "memory");
      if (ct < p.BN && n0 + ct < p.N) atomicAdd(p.ep.colsum + n0 + ct, cs_s[ct]);
    }
    if (ct == 0) TC_STAMP(4);
    // ---- epilogue: this warp owns TMEM lanes 32*(warp%4) .. +31  == output rows.
    // tcgen05.ld gives one row per thread; a per-warp smem transpose (the pipeline's stage-0 region
    // is idle once tmem_full fires) turns that into row-contiguous 128-byte global accesses.
    const int q = warp & 3;                       // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;             // two warps per quarter split the column chunks
    const Epilogue& ep = p.ep;
    const bool first = (blockIdx.z == 0);
    float* stg = reinterpret_cast<float*>(smem) + (warp - 2) * (32 * 36);
    const bool vec_ok = ((p.ldc & 3) == 0) && ((p.N & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.pre_out) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.pre_in) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.addend) & 15) == 0) &&
                        ((reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
    const int rsub = lane >> 3, csub = (lane & 7) * 4;
    const bool has_bias = (ep.bias != nullptr) && first;
    const bool add_on = ADDEND && first;
    for (int c0 = half * 32; c0 < p.BN; c0 += 64) {
      const int colv = n0 + c0 + csub;
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (vec_ok && colv < p.N) {
        // pull this chunk's global operands towards L1 before blocking on the accumulator
        if (has_bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + colv));
        if (DACT != 0 || add_on) {
#pragma unroll 1
          for (int it = 0; it < 8; it++) {
            const int gm = m0 + q * 32 + it * 4 + rsub;
            if (gm < p.M) {
              const long long idx = (long long)gm * p.ldc + colv;
              if (DACT != 0) asm volatile("prefetch.global.L1 [%0];" ::"l"(ep.pre_in + idx));
              if (add_on) asm volatile("prefetch.global.L1 [%0];" ::"l"(ep.addend + idx));
            }
          }
        }
      }
      if (c0 == half * 32) {
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        if (ct == 0) TC_STAMP(6);
      }
      {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, r);
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) =
              make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]), __uint_as_float(r[4 * j + 2]),
                          __uint_as_float(r[4 * j + 3]));
        if (wide) {
          // + the hi*lo half of the accumulator, added in the staging tile (same registers: three CTAs share an SM only
          // below 68 registers per thread)
          tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(p.BN + c0), r);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            float4 t = *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j);
            t.x += __uint_as_float(r[4 * j]); t.y += __uint_as_float(r[4 * j + 1]);
            t.z += __uint_as_float(r[4 * j + 2]); t.w += __uint_as_float(r[4 * j + 3]);
            *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) = t;
          }
        }
      }
      __syncwarp();
      if (vec_ok) {
        if (colv < p.N) {
          const int gm0 = m0 + q * 32 + rsub;
          long long idx = (long long)gm0 * p.ldc + colv;
          const long long istep = 4 * p.ldc;
          const float* sp = stg + rsub * 36 + csub;
          const float alpha = ep.alpha;
#pragma unroll 4
          for (int it = 0; it < 8; it++, idx += istep, sp += 4 * 36) {
            if (gm0 + it * 4 < p.M) {
              float4 v = *reinterpret_cast<const float4*>(sp);
              v.x = alpha * v.x + b4.x; v.y = alpha * v.y + b4.y;
              v.z = alpha * v.z + b4.z; v.w = alpha * v.w + b4.w;
              if (PRE_OUT) *reinterpret_cast<float4*>(ep.pre_out + idx) = v;
              if (ACT == NT_ACT_RELU) {
                v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
              } else if (ACT == NT_ACT_GELU) {
                v.x = gelu_f(v.x); v.y = gelu_f(v.y); v.z = gelu_f(v.z); v.w = gelu_f(v.w);
              }
              if (DACT != 0) {
                const float4 pz = __ldg(reinterpret_cast<const float4*>(ep.pre_in + idx));
                if (DACT == NT_ACT_RELU) {
                  v.x = pz.x > 0.f ? v.x : 0.f; v.y = pz.y > 0.f ? v.y : 0.f;
                  v.z = pz.z > 0.f ? v.z : 0.f; v.w = pz.w > 0.f ? v.w : 0.f;
                } else {
                  v.x *= gelu_grad_f(pz.x); v.y *= gelu_grad_f(pz.y); v.z *= gelu_grad_f(pz.z); v.w *= gelu_grad_f(pz.w);
                }
              }
              if (add_on) {
                const float4 a4 = __ldg(reinterpret_cast<const float4*>(ep.addend + idx));
                v.x += a4.x; v.y += a4.y; v.z += a4.z; v.w += a4.w;
              }
              if (ATOMIC) atomicAdd(reinterpret_cast<float4*>(p.C + idx), v);
              else *reinterpret_cast<float4*>(p.C + idx) = v;
              if (!ATOMIC && ep.out_hi) {
                uint2 h, l;
                tc_split2(v.x, v.y, h.x, l.x);
                tc_split2(v.z, v.w, h.y, l.y);
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out_hi) + idx) = h;
                *reinterpret_cast<uint2*>(reinterpret_cast<uint16_t*>(ep.out_lo) + idx) = l;
              }
            }
          }
        }
      } else {
        const int col = n0 + c0 + lane;
        if (col < p.N) {
          const float b1 = has_bias ? __ldg(ep.bias + col) : 0.f;
#pragma unroll 1
          for (int rr = 0; rr < 32; rr++) {
            const int gm = m0 + q * 32 + rr;
            if (gm >= p.M) break;
            const long long idx = (long long)gm * p.ldc + col;
            float v = ep.alpha * stg[rr * 36 + lane] + b1;
            if (PRE_OUT) ep.pre_out[idx] = v;
            if (ACT == NT_ACT_RELU) v = fmaxf(v, 0.f);
            else if (ACT == NT_ACT_GELU) v = gelu_f(v);
            if (DACT == NT_ACT_RELU) v = (__ldg(ep.pre_in + idx) > 0.f) ? v : 0.f;
            else if (DACT == NT_ACT_GELU) v *= gelu_grad_f(__ldg(ep.pre_in + idx));
            if (add_on) v += __ldg(ep.addend + idx);
            if (ATOMIC) atomicAdd(p.C + idx, v);
            else p.C[idx] = v;
            if (!ATOMIC && ep.out_hi) {
              const __nv_bfloat16 h = __float2bfloat16_rn(v);
              reinterpret_cast<__nv_bfloat16*>(ep.out_hi)[idx] = h;
              reinterpret_cast<__nv_bfloat16*>(ep.out_lo)[idx] = __float2bfloat16_rn(v - __bfloat162float(h));
            }
          }
        }
      }
      __syncwarp();
    }
    tc_fence_before();
    if (ct == 0) TC_STAMP(7);
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(tmem_cols) : "memory");
  }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int load_encode() {
  if (g_encode) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  NT_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  NT_REQUIRE(qres == cudaDriverEntryPointSuccess && fn, "cuTensorMapEncodeTiled is not available from the driver");
  g_encode = (EncodeTiledFn)fn;
  return 0;
}

// 2-D fp32 tensor map: inner dim `inner` contiguous, `outer` rows of `ld` elements; box {32, box_outer}
static int make_tmap(CUtensorMap* tm, const float* base, long long inner, long long outer, long long ld, int box_outer,
                     bool mn_major) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE,
                        mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) inner=%lld outer=%lld ld=%lld box=%d", (int)r, inner,
             outer, ld, box_outer);
  return 0;
}

// 2-D bf16 K-major tensor map: `inner` contraction elements contiguous, `outer` rows of `ld` elements; box {64, box_outer}
static int make_tmap_bf16(CUtensorMap* tm, const void* base, long long inner, long long outer, long long ld, int box_outer) {
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)outer};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(bf16) failed (%d) inner=%lld outer=%lld ld=%lld box=%d", (int)r, inner,
             outer, ld, box_outer);
  return 0;
}

// MN-major bf16 operand [k rows][mn] whose contiguous extent is a multiple of 64: view as {64, k, mn/64}, box {64, 64, chunks}
static int make_tmap_bf16_mn3d(CUtensorMap* tm, const void* base, long long mn, long long k, long long ld, int chunks) {
  cuuint64_t dims[3] = {64, (cuuint64_t)k, (cuuint64_t)(mn / 64)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, 128};
  cuuint32_t box[3] = {64, 64, (cuuint32_t)chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(bf16 3d) failed (%d) mn=%lld k=%lld ld=%lld", (int)r, mn, k, ld);
  return 0;
}

template <int ACT, int DACT, bool PRE_OUT, bool ATOMIC, bool ADDEND>
static int launch_tc(dim3 grid, size_t smem, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmAl,
                     const CUtensorMap& tmBl, const TcParams& p) {
  static bool configured = false;
  if (!configured) {
    NT_CHECK(cudaFuncSetAttribute(gemm_tc_kernel<ACT, DACT, PRE_OUT, ATOMIC, ADDEND>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
    configured = true;
  }
  NT_CHECK(nt::launch_k(gemm_tc_kernel<ACT, DACT, PRE_OUT, ATOMIC, ADDEND>, dim3(grid), dim3(TC_THREADS), smem, tmA, tmB, tmAl, tmBl, p));
  NT_LAUNCH_OK();
  return 0;
}

// N-tile: the widest of {128,96,64,32} that still yields about one wave of CTAs (small problems are
// latency-bound, so more, narrower tiles beat fewer wide ones); ties prefer exact divisors of N.
// MN-major operand whose contiguous extent is a multiple of 32: view [K][MN] as {32, K, MN/32} so one
// TMA instruction lands all 32-wide chunks of a k-block (each cp.async.bulk.tensor has a fixed cost of a few
// hundred ns in the TMA unit; four 4 KB boxes per k-block made the main loop issue-latency bound).
static int make_tmap_mn3d(CUtensorMap* tm, const float* base, long long mn, long long k, long long ld, int chunks) {
  cuuint64_t dims[3] = {32, (cuuint64_t)k, (cuuint64_t)(mn / 32)};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 4, 128};
  cuuint32_t box[3] = {32, (cuuint32_t)BK, (cuuint32_t)chunks};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NT_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed (%d) mn=%lld k=%lld ld=%lld", (int)r, mn, k, ld);
  return 0;
}

static inline bool ld_ok3(long long ld) { return (ld % 4) == 0; }

static int pick_bn(int M, int N, int splits_hint) {
  const int mt = ceil_div(M, BM) * (splits_hint < 1 ? 1 : splits_hint);
  const int cand[4] = {128, 96, 64, 32};
  int best = 32;
  long long best_score = -1;
  for (int i = 0; i < 4; i++) {
    const int bn = cand[i];
    if (bn > 32 && bn >= 2 * N) continue;
    const int tiles = mt * ceil_div(N, bn);
    const long long waste = (long long)ceil_div(N, bn) * bn - N;          // padded columns
    static int target = -1;
    if (target < 0) { const char* e = getenv("NTB200_WAVE_SMS"); target = e ? atoi(e) : g_sm_count; }
    const int fill = tiles >= (target * 3) / 4 ? target : tiles;   // >= 3/4 wave counts as full
    const long long score = (long long)fill * 1000000 - waste * 1000 + bn;
    if (score > best_score) { best_score = score; best = bn; }
  }
  return best;
}

int gemm_tc(const GemmDesc& d, int passes, bool* handled) {
  *handled = false;
  if (d.nb0 * d.nb1 != 1) return 0;                       // batched (attention) shapes stay on the SIMT path
  if (d.M < 1 || d.N < 8 || d.K < 1) return 0;
  // pre-split bf16 hi/lo K-major operands (A_b[M][lda_b], B_b[N][ldb_b]): error-compensated BF16x3 on kind::f16
  const bool bf = (d.A_hi != nullptr) && passes == 3;
  const bool a_k = bf ? !d.a_b_mn : (d.sAk == 1), a_mn = bf ? d.a_b_mn : (d.sAm == 1 && !a_k);
  const bool b_k = bf ? !d.b_b_mn : (d.sBk == 1), b_mn = bf ? d.b_b_mn : (d.sBn == 1 && !b_k);
  if (!(a_k || a_mn) || !(b_k || b_mn)) return 0;
  const long long lda = bf ? d.lda_b : (a_k ? d.sAm : d.sAk), ldb = bf ? d.ldb_b : (b_k ? d.sBn : d.sBk);
  // TMA needs 16-byte aligned bases and row pitches
  if (bf) {
    if ((lda & 7) || (ldb & 7) || ((uintptr_t)d.A_hi & 15) || ((uintptr_t)d.A_lo & 15) || ((uintptr_t)d.B_hi & 15) ||
        ((uintptr_t)d.B_lo & 15)) {
      set_error("gemm_tc: bf16 operands need 16-byte aligned bases and pitches (lda %lld ldb %lld)", lda, ldb);
      return 1;
    }
  } else if ((lda & 3) || (ldb & 3) || ((uintptr_t)d.A & 15) || ((uintptr_t)d.B & 15)) return 0;
  if (d.splitk > 1 && !d.ep.atomic) return 0;
  if (load_encode()) return 1;

  TcParams p{};
  p.M = d.M; p.N = d.N; p.K = d.K;
  if (d.splitk > 1) {
    // split-K fills the machine; take the widest tile without padded columns so A is re-read least
    p.BN = d.N % 128 == 0 ? 128 : (d.N % 96 == 0 ? 96 : (d.N % 64 == 0 ? 64 : (d.N >= 96 ? 128 : (d.N > 32 ? 64 : 32))));
    // pre-split bf16 operands, one CTA per SM with a ring: 128 x 256 tiles move 25 % fewer operand bytes per FLOP, and operand
    // delivery (not the tensor core, tools/umma_rate.cu) is what bounds these kernels — G2 weight-gradient GEMMs 9.50 -> 7.78 ms
    { static int w256 = -1; if (w256 < 0) { const char* e = getenv("NTB200_DW_BN256"); w256 = (e && e[0] == '0') ? 0 : 1; }
      // (a ragged last tile is accepted while it wastes < 15 %: N = 1152 -> 5 tiles instead of 9)
      if (bf && w256 && d.N >= 256 && (long long)ceil_div(d.N, 256) * 256 * 100 <= 115LL * d.N) p.BN = 256; }
  } else {
    p.BN = pick_bn(d.M, d.N, 1);
  }
  p.a_mn = a_mn; p.b_mn = b_mn;
  p.passes = passes;
  p.bf16 = bf ? 1 : 0;
  p.wide = 0;
  if (bf && b_mn && (p.BN % 64)) p.BN = d.N > 64 ? 128 : 64;       // MN-major bf16 tiles come in 64-wide chunks
  p.ldc = d.ldc;
  p.C = d.C;
  p.ep = d.ep;
  { const char* e = getenv("NTB200_DBG_MMA"); p.dbg_mma = e ? atoi(e) : 4; }
  { const char* e = getenv("NTB200_DBG_SKIP_A"); p.dbg_skip_a = (e && e[0] == '1') ? 1 : 0; }
  if (const char* e = getenv("NTB200_BN")) { int v = atoi(e); if (v == 32 || v == 64 || v == 96 || v == 128 || (bf && v == 256 && d.N >= 256)) p.BN = v; }
  { static int w = -1; if (w < 0) { const char* e = getenv("NTB200_GEMM_WIDE"); w = (e && e[0] == '0') ? 0 : 1; }
    p.wide = (bf && w && d.ep.atomic && 2 * p.BN <= 256) ? 1 : 0; }
  const int total_kb = ceil_div(d.K, bf ? 2 * BK : BK);
  int splits = 1;
  if (d.splitk > 1) {
    // split the long contraction (dW: K = batch rows) so the grid covers the SMs; >= 4 k-blocks per split
    const int tiles = ceil_div(d.N, p.BN) * ceil_div(d.M, BM);
    splits = ceil_div(g_sm_count, tiles);
    // the deep-ring bf16 configuration runs ONE CTA per SM: a grid a little above the SM count costs a whole second
    // wave (ncu: 162 CTAs = 148 + 14 ran at 33 % SM throughput), so round the split count DOWN to stay within one wave
    if (bf && tiles <= g_sm_count) { splits = g_sm_count / tiles; if (splits < 1) splits = 1; }
    const int maxs = total_kb / 4 < 1 ? 1 : total_kb / 4;
    if (splits > maxs) splits = maxs;
  }
  p.kblocks_per_split = ceil_div(total_kb, splits);
  splits = ceil_div(total_kb, p.kblocks_per_split);
  const size_t stage = (size_t)(BM * BK * 4 + p.BN * BK * 4) * (passes == 3 ? 2 : 1);
  int stages = (int)((200 * 1024) / stage);
  if (stages > 6) stages = 6;
  if (stages > p.kblocks_per_split) stages = p.kblocks_per_split < 2 ? 2 : p.kblocks_per_split;
  // single-pass tiles are small (32 KB/stage): with two stages three CTAs share an SM and one CTA's
  // prologue / epilogue hides behind another's main loop (measured 218 -> 115 us on 50176x384x1152)
  if (passes == 1 && (long long)ceil_div(d.N, p.BN) * ceil_div(d.M, BM) * splits >= 2LL * g_sm_count && stages > 2) stages = 2;
  // pre-split bf16 tiles need no converter hand-off: without split-K (forward / dX, <= 64 k-blocks) one 64 KB stage
  // lets three CTAs share an SM, so loads, MMAs and epilogues of different tiles overlap (measured on V2: forward /
  // dX GEMMs 6.5 -> 4.6 ms; G2 dX 9.0 -> 8.0 ms); the long split-K weight-gradient contraction keeps the deep ring
  { static int lim = -1; if (lim < 0) { const char* e = getenv("NTB200_BF16_1STAGE_KB"); lim = e ? atoi(e) : 64; }
    if (bf && splits == 1 && p.kblocks_per_split <= lim) stages = 1; }
  if (const char* e = getenv("NTB200_STAGES")) { int v = atoi(e); if (v >= 1 && v <= stages) stages = v; }
  p.stages = stages;
  const size_t smem = stages * stage + (3 * stages + 2) * 8 + 16 + 128 * 4 + 1024;

  CUtensorMap tmA, tmB, tmAl, tmBl;
  if (bf) {
    // K-major: [rows = M or N][K contiguous], box {64 k, rows of the tile}; MN-major: [K rows][M or N contiguous], box {64, 64}
    const bool three_d = getenv("NTB200_TMA_2D") == nullptr;
    p.a_3d = (a_mn && three_d && d.M % 64 == 0) ? 1 : 0;
    p.b_3d = (b_mn && three_d && d.N % 64 == 0) ? 1 : 0;
    if (p.a_3d ? (make_tmap_bf16_mn3d(&tmA, d.A_hi, d.M, d.K, lda, BM / 64) || make_tmap_bf16_mn3d(&tmAl, d.A_lo, d.M, d.K, lda, BM / 64))
        : a_mn ? (make_tmap_bf16(&tmA, d.A_hi, d.M, d.K, lda, 64) || make_tmap_bf16(&tmAl, d.A_lo, d.M, d.K, lda, 64))
               : (make_tmap_bf16(&tmA, d.A_hi, d.K, d.M, lda, BM) || make_tmap_bf16(&tmAl, d.A_lo, d.K, d.M, lda, BM)))
      return 1;
    if (p.b_3d ? (make_tmap_bf16_mn3d(&tmB, d.B_hi, d.N, d.K, ldb, p.BN / 64) || make_tmap_bf16_mn3d(&tmBl, d.B_lo, d.N, d.K, ldb, p.BN / 64))
        : b_mn ? (make_tmap_bf16(&tmB, d.B_hi, d.N, d.K, ldb, 64) || make_tmap_bf16(&tmBl, d.B_lo, d.N, d.K, ldb, 64))
               : (make_tmap_bf16(&tmB, d.B_hi, d.K, d.N, ldb, p.BN) || make_tmap_bf16(&tmBl, d.B_lo, d.K, d.N, ldb, p.BN)))
      return 1;
  }
  const bool use3d = getenv("NTB200_TMA_2D") == nullptr;
  if (!bf) {
    p.a_3d = (a_mn && use3d && d.M % 32 == 0 && ld_ok3(lda)) ? 1 : 0;
    p.b_3d = (b_mn && use3d && d.N % 32 == 0 && ld_ok3(ldb)) ? 1 : 0;
  }
  if (bf) {}
  else if (a_k) { if (make_tmap(&tmA, d.A, d.K, d.M, lda, BM, false)) return 1; }
  else if (p.a_3d) { if (make_tmap_mn3d(&tmA, d.A, d.M, d.K, lda, BM / 32)) return 1; }
  else     { if (make_tmap(&tmA, d.A, d.M, d.K, lda, BK, true)) return 1; }
  if (bf) {}
  else if (b_k) { if (make_tmap(&tmB, d.B, d.K, d.N, ldb, p.BN, false)) return 1; }
  else if (p.b_3d) { if (make_tmap_mn3d(&tmB, d.B, d.N, d.K, ldb, p.BN / 32)) return 1; }
  else     { if (make_tmap(&tmB, d.B, d.N, d.K, ldb, BK, true)) return 1; }

  if (!bf) { tmAl = tmA; tmBl = tmB; }
  dim3 grid(ceil_div(d.N, p.BN), ceil_div(d.M, BM), splits);
  const Epilogue& e = d.ep;
  int rc = -1;
#define NT_TC_CASE(A_, D_, P_, T_, X_)                                                                   \
  if (rc < 0 && e.act == A_ && e.dact == D_ && (e.pre_out != nullptr) == P_ && (e.atomic != 0) == T_ &&   \
      (e.addend != nullptr) == X_)                                                                        \
    rc = launch_tc<A_, D_, P_, T_, X_>(grid, smem, tmA, tmB, tmAl, tmBl, p);
  NT_TC_CASE(0, 0, false, false, false)   // y = xW + b ; dx = dy W^T
  NT_TC_CASE(0, 0, false, false, true)    // + residual
  NT_TC_CASE(NT_ACT_GELU, 0, true, false, false)
  NT_TC_CASE(NT_ACT_RELU, 0, true, false, false)
  NT_TC_CASE(NT_ACT_GELU, 0, true, false, true)
  NT_TC_CASE(NT_ACT_RELU, 0, true, false, true)
  NT_TC_CASE(0, NT_ACT_GELU, false, false, false)
  NT_TC_CASE(0, NT_ACT_RELU, false, false, false)
  NT_TC_CASE(0, NT_ACT_GELU, false, false, true)
  NT_TC_CASE(0, NT_ACT_RELU, false, false, true)
  NT_TC_CASE(0, 0, false, true, false)    // dW += x^T dy (split-K, red.add)
#undef NT_TC_CASE
  if (rc < 0) return 0;                   // epilogue combination without a tensor-core instance -> SIMT
  if (rc) return rc;
  if (d.ep.colsum && d.colsum_done) *d.colsum_done = (b_mn != 0);
  *handled = true;
  return 0;
}

}  // namespace nt
