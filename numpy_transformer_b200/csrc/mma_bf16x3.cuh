// Warp-level building blocks of the error-compensated BF16x3 mma.sync path (shared by the fused ViT-block kernels,
// vit_block.cu, and the attention kernels, attention_bf16.cu): fp32 -> (hi, lo) bf16 split, split operand arrays in
// shared memory, ldmatrix fragment loads, m16n8k16 / m16n8k8 MMAs and the small tile GEMMs built from them.
#pragma once
#include "common.cuh"
#include <cuda_bf16.h>

namespace nt {
namespace vb {

constexpr int ROWS = 64;      // padded tokens / tile height
constexpr int LDP = 72;       // pitch (elements) of [*][64] score tiles: 144 bytes = 16 (mod 128)

// ---------------------------------------------------------------- primitives
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}
__device__ __forceinline__ void split1(float a, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(a);
  lo = __float2bfloat16_rn(a - __bfloat162float(hi));
}

// a [64][ld] bf16 operand array stored as two planes (hi, then lo)
struct SArr {
  char* ptr;        // generic pointer to the hi plane
  uint32_t addr;    // shared-window address of the hi plane
  uint32_t ldb;     // row pitch in bytes
  uint32_t lo;      // byte offset hi -> lo plane
};
__device__ __forceinline__ SArr make_arr(char* p, int ld_elems, int rows = ROWS) {
  SArr a;
  a.ptr = p;
  a.addr = (uint32_t)__cvta_generic_to_shared(p);
  a.ldb = ld_elems * 2;
  a.lo = rows * ld_elems * 2;
  return a;
}
__device__ __forceinline__ void st2(const SArr& a, int row, int col, float v0, float v1) {
  uint32_t hi, lo;
  split2(v0, v1, hi, lo);
  char* p = a.ptr + row * a.ldb + col * 2;
  *reinterpret_cast<uint32_t*>(p) = hi;
  *reinterpret_cast<uint32_t*>(p + a.lo) = lo;
}
__device__ __forceinline__ void st1(const SArr& a, int row, int col, float v) {
  __nv_bfloat16 hi, lo;
  split1(v, hi, lo);
  char* p = a.ptr + row * a.ldb + col * 2;
  *reinterpret_cast<__nv_bfloat16*>(p) = hi;
  *reinterpret_cast<__nv_bfloat16*>(p + a.lo) = lo;
}

__device__ __forceinline__ void ldsm4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void ldsm2t(uint32_t (&r)[2], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}
__device__ __forceinline__ void mma16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a0), "r"(a1), "r"(b0));
}
__device__ __forceinline__ void red2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

template <int MT, int NT>
__device__ __forceinline__ void zero_acc(float (&acc)[MT][NT][4]) {
#pragma unroll
  for (int m = 0; m < MT; m++)
#pragma unroll
    for (int n = 0; n < NT; n++) acc[m][n][0] = acc[m][n][1] = acc[m][n][2] = acc[m][n][3] = 0.f;
}

// first k-step of a weight-fragment stream: issued a whole phase ahead of the GEMM that consumes it, so the L2
// round trip overlaps the LayerNorm / epilogue in between instead of stalling the first MMAs
template <int NTW>
__device__ __forceinline__ void wf_first(uint4 (&b0)[NTW], const uint4* __restrict__ wf) {
#pragma unroll
  for (int nt = 0; nt < NTW; nt++) b0[nt] = __ldg(wf + nt * 32);
}

// acc[mt][nt] += A(rows, k) * W, A row-major in shared memory (hi/lo planes), W as pre-split B fragments in
// global memory ([k16 step][n8 tile][lane] uint4 = {b0.hi, b1.hi, b0.lo, b1.lo}).  KS k-steps, fully unrolled:
// B fragments run PF k-steps ahead in a register ring, A fragments one k-step ahead.
// a_addr(ks): shared address (hi plane) of this lane's ldmatrix row for m-tile 0 at k-step ks.
template <int MT, int NTW, int KS, int PF, class AF>
__device__ __forceinline__ void gemm_wf(float (&acc)[MT][NTW][4], AF a_addr, uint32_t a_lo, uint32_t a_mt_stride,
                                        const uint4* __restrict__ wf, int ks_stride, const uint4 (&b0)[NTW]) {
  uint4 b[PF + 1][NTW];
#pragma unroll
  for (int nt = 0; nt < NTW; nt++) b[0][nt] = b0[nt];
#pragma unroll
  for (int p = 1; p <= PF; p++)
    if (p < KS) {
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) b[p][nt] = __ldg(wf + (size_t)p * ks_stride + nt * 32);
    }
  uint32_t ah[2][MT][4], al[2][MT][4];
  {
    const uint32_t a0 = a_addr(0);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      ldsm4(ah[0][mt], a0 + mt * a_mt_stride);
      ldsm4(al[0][mt], a0 + mt * a_mt_stride + a_lo);
    }
  }
#pragma unroll
  for (int ks = 0; ks < KS; ks++) {
    const int ab = ks & 1, slot = ks % (PF + 1);
    if (ks + 1 < KS) {
      const uint32_t a1 = a_addr(ks + 1);
#pragma unroll
      for (int mt = 0; mt < MT; mt++) {
        ldsm4(ah[ab ^ 1][mt], a1 + mt * a_mt_stride);
        ldsm4(al[ab ^ 1][mt], a1 + mt * a_mt_stride + a_lo);
      }
    }
    // pass-major: consecutive MMAs hit different accumulators (no back-to-back RAW on the tensor pipe)
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], al[ab][mt], b[slot][nt].x, b[slot][nt].y);
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], ah[ab][mt], b[slot][nt].z, b[slot][nt].w);
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], ah[ab][mt], b[slot][nt].x, b[slot][nt].y);
    if (ks + PF + 1 < KS) {
#pragma unroll
      for (int nt = 0; nt < NTW; nt++) b[slot][nt] = __ldg(wf + (size_t)(ks + PF + 1) * ks_stride + nt * 32);
    }
  }
}

// acc[mt][nt] += A^T B with both operands stored [k][*] in shared memory (contraction over the row index = tokens):
// A(m, k) = SA[k][m0 + m], B(k, n) = SB[k][n0 + n]; fragments via ldmatrix.trans.
template <int MT, int NTW>
__device__ __forceinline__ void gemm_tt(float (&acc)[MT][NTW][4], const SArr& A, int m0, const SArr& B, int n0,
                                        int ksteps, int lane) {
  const uint32_t a_lane = A.addr + ((lane & 7) + ((lane >> 4) & 1) * 8) * A.ldb + (m0 + ((lane >> 3) & 1) * 8) * 2;
  const uint32_t b_lane = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + (n0 + (lane >> 4) * 8) * 2;
  const uint32_t b_lane2 = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + n0 * 2;
#pragma unroll
  for (int ks = 0; ks < ksteps; ks++) {
    uint32_t ah[MT][4], al[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      ldsm4t(ah[mt], a_lane + ks * 16 * A.ldb + mt * 32);
      ldsm4t(al[mt], a_lane + ks * 16 * A.ldb + mt * 32 + A.lo);
    }
    uint32_t bh[NTW][2], bl[NTW][2];
#pragma unroll
    for (int nt = 0; nt + 1 < NTW; nt += 2) {
      uint32_t h4[4], l4[4];
      ldsm4t(h4, b_lane + ks * 16 * B.ldb + nt * 16);
      ldsm4t(l4, b_lane + ks * 16 * B.ldb + nt * 16 + B.lo);
      bh[nt][0] = h4[0]; bh[nt][1] = h4[1]; bh[nt + 1][0] = h4[2]; bh[nt + 1][1] = h4[3];
      bl[nt][0] = l4[0]; bl[nt][1] = l4[1]; bl[nt + 1][0] = l4[2]; bl[nt + 1][1] = l4[3];
    }
    if (NTW & 1) {
      ldsm2t(bh[NTW - 1], b_lane2 + ks * 16 * B.ldb + (NTW - 1) * 16);
      ldsm2t(bl[NTW - 1], b_lane2 + ks * 16 * B.ldb + (NTW - 1) * 16 + B.lo);
    }
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], al[mt], bh[nt][0], bh[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], ah[mt], bl[nt][0], bl[nt][1]);
#pragma unroll
    for (int nt = 0; nt < NTW; nt++)
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16(acc[mt][nt], ah[mt], bh[nt][0], bh[nt][1]);
  }
}

// s[nt] += A(16 rows at a_row0, k over [kcol, kcol+DH)) * Bt where B is stored [n][k] ("nk": rows = output
// columns).  8 output tiles (64 columns) in pairs; DH = 16*a + 8*b.  MMAs are issued pass-major so that
// consecutive instructions never share an accumulator.
template <int DH>
__device__ __forceinline__ void gemm_rowk_nk(float (&s)[8][4], const SArr& A, int a_row0, const SArr& B, int kcol,
                                             int npairs, int lane) {
  const uint32_t a_lane = A.addr + (a_row0 + (lane & 15)) * A.ldb + (kcol + (lane >> 4) * 8) * 2;
  const uint32_t b_lane = B.addr + ((lane & 7) + (lane >> 4) * 8) * B.ldb + (kcol + ((lane >> 3) & 1) * 8) * 2;
#pragma unroll
  for (int kk = 0; kk < DH / 16; kk++) {
    uint32_t ah[4], al[4], bh[4][4], bl[4][4];
    ldsm4(ah, a_lane + kk * 32);
    ldsm4(al, a_lane + kk * 32 + A.lo);
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) {
        ldsm4(bh[p], b_lane + p * 16 * B.ldb + kk * 32);
        ldsm4(bl[p], b_lane + p * 16 * B.ldb + kk * 32 + B.lo);
      }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma16(s[2 * p], al, bh[p][0], bh[p][1]); mma16(s[2 * p + 1], al, bh[p][2], bh[p][3]); }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma16(s[2 * p], ah, bl[p][0], bl[p][1]); mma16(s[2 * p + 1], ah, bl[p][2], bl[p][3]); }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma16(s[2 * p], ah, bh[p][0], bh[p][1]); mma16(s[2 * p + 1], ah, bh[p][2], bh[p][3]); }
  }
  if (DH % 16) {
    constexpr int k0 = (DH / 16) * 16;
    const uint32_t a2 = A.addr + (a_row0 + (lane & 15)) * A.ldb + (kcol + k0) * 2;
    const uint32_t b2 = B.addr + (lane & 15) * B.ldb + (kcol + k0) * 2;
    uint32_t ah[2], al[2], bh[4][2], bl[4][2];
    ldsm2(ah, a2);
    ldsm2(al, a2 + A.lo);
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) {
        ldsm2(bh[p], b2 + p * 16 * B.ldb);
        ldsm2(bl[p], b2 + p * 16 * B.ldb + B.lo);
      }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma8(s[2 * p], al[0], al[1], bh[p][0]); mma8(s[2 * p + 1], al[0], al[1], bh[p][1]); }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma8(s[2 * p], ah[0], ah[1], bl[p][0]); mma8(s[2 * p + 1], ah[0], ah[1], bl[p][1]); }
#pragma unroll
    for (int p = 0; p < 4; p++)
      if (p < npairs) { mma8(s[2 * p], ah[0], ah[1], bh[p][0]); mma8(s[2 * p + 1], ah[0], ah[1], bh[p][1]); }
  }
}

// B fragments of NDT n8 tiles at columns ncol.. of a [k][n] ("kn") operand for one k16 step
template <int NDT>
__device__ __forceinline__ void load_b_kn(uint32_t (&bh)[NDT][2], uint32_t (&bl)[NDT][2], const SArr& B, uint32_t b_lane,
                                          uint32_t b_lane2, int kk) {
#pragma unroll
  for (int nd = 0; nd + 1 < NDT; nd += 2) {
    uint32_t h4[4], l4[4];
    ldsm4t(h4, b_lane + kk * 16 * B.ldb + nd * 16);
    ldsm4t(l4, b_lane + kk * 16 * B.ldb + nd * 16 + B.lo);
    bh[nd][0] = h4[0]; bh[nd][1] = h4[1]; bh[nd + 1][0] = h4[2]; bh[nd + 1][1] = h4[3];
    bl[nd][0] = l4[0]; bl[nd][1] = l4[1]; bl[nd + 1][0] = l4[2]; bl[nd + 1][1] = l4[3];
  }
  if (NDT & 1) {
    ldsm2t(bh[NDT - 1], b_lane2 + kk * 16 * B.ldb + (NDT - 1) * 16);
    ldsm2t(bl[NDT - 1], b_lane2 + kk * 16 * B.ldb + (NDT - 1) * 16 + B.lo);
  }
}
// the three passes of one k16 step on NDT output tiles; the small cross terms accumulate in `os`, hi*hi in `o`,
// so the dependent chain per accumulator is one MMA per k-step per set
template <int NDT>
__device__ __forceinline__ void mma_passes(float (&o)[NDT][4], float (&os)[NDT][4], const uint32_t (&ah)[4],
                                           const uint32_t (&al)[4], const uint32_t (&bh)[NDT][2],
                                           const uint32_t (&bl)[NDT][2]) {
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) mma16(os[nd], al, bh[nd][0], bh[nd][1]);
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) mma16(o[nd], ah, bh[nd][0], bh[nd][1]);
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) mma16(os[nd], ah, bl[nd][0], bl[nd][1]);
}

// o[nd] += A(16 rows at a_row0, k = tokens 0..16*ksteps) * B where B is stored [k][n] ("kn"), columns
// ncol .. ncol+DH.  A row-major [row][token].
template <int DH>
__device__ __forceinline__ void gemm_rowk_kn(float (&o)[DH / 8][4], const SArr& A, int a_row0, const SArr& B, int ncol,
                                             int ksteps, int lane) {
  constexpr int NDT = DH / 8;
  const uint32_t a_lane = A.addr + (a_row0 + (lane & 15)) * A.ldb + ((lane >> 4) * 8) * 2;
  const uint32_t b_lane = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + (ncol + (lane >> 4) * 8) * 2;
  const uint32_t b_lane2 = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + ncol * 2;
  float os[NDT][4];
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) os[nd][0] = os[nd][1] = os[nd][2] = os[nd][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < ksteps; kk++) {
    uint32_t ah[4], al[4], bh[NDT][2], bl[NDT][2];
    ldsm4(ah, a_lane + kk * 32);
    ldsm4(al, a_lane + kk * 32 + A.lo);
    load_b_kn<NDT>(bh, bl, B, b_lane, b_lane2, kk);
    mma_passes<NDT>(o, os, ah, al, bh, bl);
  }
#pragma unroll
  for (int nd = 0; nd < NDT; nd++)
#pragma unroll
    for (int c = 0; c < 4; c++) o[nd][c] += os[nd][c];
}

// o[nd] += P * B with P (16 rows x 16*ksteps k) still in the accumulator-fragment layout of the MMA that produced it:
// the C fragments of two adjacent n8 tiles ARE the A fragment of one k16 step (rows g / g+8, columns 2t.. and 2t+8..),
// so the probabilities never travel through shared memory.  Same split (hi, lo) values as st2 + ldmatrix would give.
template <int DH>
__device__ __forceinline__ void gemm_regA_kn(float (&o)[DH / 8][4], const float (&p)[8][4], const SArr& B, int ncol,
                                             int ksteps, int lane) {
  constexpr int NDT = DH / 8;
  const uint32_t b_lane = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + (ncol + (lane >> 4) * 8) * 2;
  const uint32_t b_lane2 = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + ncol * 2;
  float os[NDT][4];
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) os[nd][0] = os[nd][1] = os[nd][2] = os[nd][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < 4; kk++)
    if (kk < ksteps) {
      uint32_t ah[4], al[4], bh[NDT][2], bl[NDT][2];
      split2(p[2 * kk][0], p[2 * kk][1], ah[0], al[0]);
      split2(p[2 * kk][2], p[2 * kk][3], ah[1], al[1]);
      split2(p[2 * kk + 1][0], p[2 * kk + 1][1], ah[2], al[2]);
      split2(p[2 * kk + 1][2], p[2 * kk + 1][3], ah[3], al[3]);
      load_b_kn<NDT>(bh, bl, B, b_lane, b_lane2, kk);
      mma_passes<NDT>(o, os, ah, al, bh, bl);
    }
#pragma unroll
  for (int nd = 0; nd < NDT; nd++)
#pragma unroll
    for (int c = 0; c < 4; c++) o[nd][c] += os[nd][c];
}

// o[nd] += A^T(16 output rows = columns m0.. of SA, k = rows of SA) * B ("kn"): both stored [k][*]
template <int DH>
__device__ __forceinline__ void gemm_colk_kn(float (&o)[DH / 8][4], const SArr& A, int m0, const SArr& B, int ncol,
                                             int ksteps, int lane) {
  constexpr int NDT = DH / 8;
  const uint32_t a_lane = A.addr + ((lane & 7) + ((lane >> 4) & 1) * 8) * A.ldb + (m0 + ((lane >> 3) & 1) * 8) * 2;
  const uint32_t b_lane = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + (ncol + (lane >> 4) * 8) * 2;
  const uint32_t b_lane2 = B.addr + ((lane & 7) + ((lane >> 3) & 1) * 8) * B.ldb + ncol * 2;
  float os[NDT][4];
#pragma unroll
  for (int nd = 0; nd < NDT; nd++) os[nd][0] = os[nd][1] = os[nd][2] = os[nd][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < ksteps; kk++) {
    uint32_t ah[4], al[4], bh[NDT][2], bl[NDT][2];
    ldsm4t(ah, a_lane + kk * 16 * A.ldb);
    ldsm4t(al, a_lane + kk * 16 * A.ldb + A.lo);
    load_b_kn<NDT>(bh, bl, B, b_lane, b_lane2, kk);
    mma_passes<NDT>(o, os, ah, al, bh, bl);
  }
#pragma unroll
  for (int nd = 0; nd < NDT; nd++)
#pragma unroll
    for (int c = 0; c < 4; c++) o[nd][c] += os[nd][c];
}

}  // namespace vb
}  // namespace nt
