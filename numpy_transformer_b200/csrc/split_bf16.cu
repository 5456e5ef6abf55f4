// bf16 hi/lo split of fp32 matrices for the error-compensated BF16x3 tensor-core GEMM (gemm_tc.cu, bf16 mode):
//   hi = bf16(x), lo = bf16(x - hi)   =>   x = hi + lo up to ~2^-17 relative,
// and a product a*b is accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in fp32 (TMEM).  The split is done once per
// operand in HBM (optionally transposed, so that every GEMM of the Linear layer sees K-major operands:
// net/fullconnect.py:103-105 forward, :114 dX, :118-121 dW) instead of per k-block inside the GEMM, where the
// fp32 -> (hi, lo) conversion made the 3xTF32 kernel shared-memory-bandwidth bound.
#include "common.cuh"
#include <cuda_bf16.h>

namespace nt {

__device__ __forceinline__ void split2_bf16(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = *reinterpret_cast<uint32_t*>(&l);
}

// out[r][c] planes with pitch ld_out (elements); cols % 4 == 0, ld_out % 4 == 0
__global__ void __launch_bounds__(256) split_rows_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, long long rows, int cols,
                                                         int ld_out) {
  pdl_prologue();
  const int c4 = cols >> 2;
  const long long total = rows * c4;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long r = e / c4;
    const int c = (int)(e - r * c4) * 4;
    const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * cols + c));
    uint2 h, l;
    split2_bf16(v.x, v.y, h.x, l.x);
    split2_bf16(v.z, v.w, h.y, l.y);
    *reinterpret_cast<uint2*>(hi + r * ld_out + c) = h;
    *reinterpret_cast<uint2*>(lo + r * ld_out + c) = l;
  }
}

// transposing split: x [rows][cols] fp32 -> hiT / loT [cols][ld_out] (ld_out >= rows, % 4 == 0); 64 x 64 tiles
// BOTH != 0: the same pass also writes the untransposed planes hi / lo [rows][cols] (cols % 4 == 0), so a tensor
// that feeds one GEMM as-is and another transposed (x: forward and dW; dy: dX and dW) is read from HBM once.
template <int BOTH>
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ hiT,
                                                              __nv_bfloat16* __restrict__ loT, int rows, int cols, int ld_out,
                                                              __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo) {
  pdl_prologue();
  __shared__ float tile[64][65];
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const int tid = threadIdx.x;
  const bool vec = (cols & 3) == 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int e = tid + 256 * i;             // 1024 float4 slots: row = e / 16, chunk = e % 16
    const int r = e >> 4, c = (e & 15) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r0 + r < rows) {
      const float* src = x + (long long)(r0 + r) * cols + c0 + c;
      if (vec && c0 + c + 3 < cols) v = __ldg(reinterpret_cast<const float4*>(src));
      else {
        if (c0 + c < cols) v.x = __ldg(src);
        if (c0 + c + 1 < cols) v.y = __ldg(src + 1);
        if (c0 + c + 2 < cols) v.z = __ldg(src + 2);
        if (c0 + c + 3 < cols) v.w = __ldg(src + 3);
      }
    }
    tile[r][c] = v.x; tile[r][c + 1] = v.y; tile[r][c + 2] = v.z; tile[r][c + 3] = v.w;
    if (BOTH && r0 + r < rows && c0 + c < cols) {
      uint2 h, l;
      split2_bf16(v.x, v.y, h.x, l.x);
      split2_bf16(v.z, v.w, h.y, l.y);
      const long long o = (long long)(r0 + r) * cols + c0 + c;
      *reinterpret_cast<uint2*>(hi + o) = h;
      *reinterpret_cast<uint2*>(lo + o) = l;
    }
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int e = tid + 256 * i;             // output row (= input column) e / 16, 4-element chunk e % 16 along input rows
    const int oc = e >> 4, q = (e & 15) * 4;
    if (c0 + oc < cols && r0 + q < ld_out) {
      uint2 h, l;
      split2_bf16(tile[q][oc], tile[q + 1][oc], h.x, l.x);
      split2_bf16(tile[q + 2][oc], tile[q + 3][oc], h.y, l.y);
      const long long o = (long long)(c0 + oc) * ld_out + r0 + q;
      *reinterpret_cast<uint2*>(hiT + o) = h;       // rows beyond `rows` were zero-filled in the tile
      *reinterpret_cast<uint2*>(loT + o) = l;
    }
  }
}

}  // namespace nt
using namespace nt;

extern "C" {

int nt_split_bf16(const float* x, void* hi, void* lo, int rows, int cols, int ld_out, int transpose) {
  NT_ENTRY();
  NT_REQUIRE(rows >= 0 && cols > 0 && (ld_out & 3) == 0, "nt_split_bf16: rows %d cols %d ld_out %d (ld_out %% 4 != 0)", rows, cols, ld_out);
  if (rows == 0) return 0;
  if (!transpose) {
    NT_REQUIRE((cols & 3) == 0 && ld_out >= cols, "nt_split_bf16: cols %d must be a multiple of 4 and <= ld_out %d", cols, ld_out);
    const long long total = (long long)rows * (cols >> 2);
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    NT_CHECK(launch_k(split_rows_kernel, dim3(blocks), dim3(256), 0, x, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, (long long)rows,
                      cols, ld_out));
  } else {
    NT_REQUIRE(ld_out >= rows, "nt_split_bf16: transposed pitch %d < rows %d", ld_out, rows);
    NT_CHECK(launch_k(split_transpose_kernel<0>, dim3(ceil_div(cols, 64), ceil_div(rows, 64)), dim3(256), 0, x,
                      (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, cols, ld_out, (__nv_bfloat16*)nullptr,
                      (__nv_bfloat16*)nullptr));
  }
  NT_LAUNCH_OK();
  return 0;
}

int nt_split_bf16_both(const float* x, void* hi, void* lo, void* hi_t, void* lo_t, int rows, int cols, int ld_t) {
  NT_ENTRY();
  NT_REQUIRE(rows >= 0 && cols > 0 && (cols & 3) == 0 && (ld_t & 3) == 0 && ld_t >= rows,
             "nt_split_bf16_both: rows %d cols %d (cols %% 4) ld_t %d", rows, cols, ld_t);
  if (rows == 0) return 0;
  NT_CHECK(launch_k(split_transpose_kernel<1>, dim3(ceil_div(cols, 64), ceil_div(rows, 64)), dim3(256), 0, x,
                    (__nv_bfloat16*)hi_t, (__nv_bfloat16*)lo_t, rows, cols, ld_t, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo));
  NT_LAUNCH_OK();
  return 0;
}

}
