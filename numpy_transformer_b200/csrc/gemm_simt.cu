// fp32 SIMT GEMM with generic strides, two-level batching, split-K and the shared epilogue.
// Role: exact-fp32 engine (gemm mode 0), fallback for shapes the tcgen05 path cannot take
// (unaligned leading dimensions, tiny attention tiles), and on-device cross-check for it.
#include "common.cuh"

namespace nt {

template <int BM, int BN, int BK>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmDesc d) {
  pdl_prologue();
  constexpr int TM = BM / 16, TN = BN / 16;
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int tid = threadIdx.x;
  const int tx = tid % 16, ty = tid / 16;
  int z = blockIdx.z;
  const int ks = z % d.splitk;
  z /= d.splitk;
  const int z0 = z / d.nb1, z1 = z % d.nb1;
  const float* __restrict__ A = d.A + z0 * d.bA0 + z1 * d.bA1;
  const float* __restrict__ B = d.B + z0 * d.bB0 + z1 * d.bB1;
  float* __restrict__ C = d.C + z0 * d.bC0 + z1 * d.bC1;
  const long long coff = z0 * d.bC0 + z1 * d.bC1;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  // K range of this split
  const int kchunk = ((d.K + d.splitk - 1) / d.splitk + BK - 1) / BK * BK;
  const int kbeg = ks * kchunk;
  const int kend = min(d.K, kbeg + kchunk);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; i++)
#pragma unroll
    for (int j = 0; j < TN; j++) acc[i][j] = 0.f;

  const bool a_kcontig = (d.sAk == 1);
  const bool b_ncontig = (d.sBn == 1);

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- A tile (BM x BK)
    if (a_kcontig) {
      const int kk = tid % BK;
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; i++) {
        int mm = tid / BK + i * (256 / BK);
        int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < d.M && gk < kend) ? __ldg(A + gm * d.sAm + gk) : 0.f;
      }
    } else {
      const int mm = tid % BM;
#pragma unroll
      for (int i = 0; i < (BM * BK) / 256; i++) {
        int kk = tid / BM + i * (256 / BM);
        int gm = m0 + mm, gk = k0 + kk;
        As[kk][mm] = (gm < d.M && gk < kend) ? __ldg(A + gm * d.sAm + gk * d.sAk) : 0.f;
      }
    }
    // ---- B tile (BK x BN)
    if (b_ncontig) {
      const int nn = tid % BN;
#pragma unroll
      for (int i = 0; i < (BN * BK) / 256; i++) {
        int kk = tid / BN + i * (256 / BN);
        int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < d.N && gk < kend) ? __ldg(B + gk * d.sBk + gn) : 0.f;
      }
    } else {
      const int kk = tid % BK;
#pragma unroll
      for (int i = 0; i < (BN * BK) / 256; i++) {
        int nn = tid / BK + i * (256 / BK);
        int gn = n0 + nn, gk = k0 + kk;
        Bs[kk][nn] = (gn < d.N && gk < kend) ? __ldg(B + gk * d.sBk + gn * d.sBn) : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; kk++) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i++) a[i] = As[kk][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < TN; j++) b[j] = Bs[kk][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < TM; i++)
#pragma unroll
        for (int j = 0; j < TN; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  const Epilogue& ep = d.ep;
  const bool first = (ks == 0);
#pragma unroll
  for (int i = 0; i < TM; i++) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= d.M) continue;
#pragma unroll
    for (int j = 0; j < TN; j++) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= d.N) continue;
      const long long idx = (long long)gm * d.ldc + gn;
      float v = ep.alpha * acc[i][j];
      if (ep.bias && first) v += __ldg(ep.bias + gn);
      if (ep.pre_out) ep.pre_out[coff + idx] = v;
      if (ep.act == NT_ACT_RELU) v = v < 0.f ? 0.f : v;
      else if (ep.act == NT_ACT_GELU) v = gelu_f(v);
      if (ep.dact == NT_ACT_RELU) v = (__ldg(ep.pre_in + coff + idx) > 0.f) ? v : 0.f;
      else if (ep.dact == NT_ACT_GELU) v *= gelu_grad_f(__ldg(ep.pre_in + coff + idx));
      if (ep.addend && first) v += __ldg(ep.addend + coff + idx);
      if (ep.atomic) atomicAdd(C + idx, v);
      else C[idx] = v;
    }
  }
}

int gemm_simt(const GemmDesc& d) {
  if (d.M <= 0 || d.N <= 0) return 0;
  NT_REQUIRE(d.splitk == 1 || (d.ep.atomic && d.ep.act == 0 && d.ep.dact == 0 && !d.ep.pre_out),
             "gemm_simt: split-K needs a pure atomic-accumulate epilogue");
  const int nb = d.nb0 * d.nb1;
  if (d.N <= 32 && d.M >= 64) {
    dim3 grid(ceil_div(d.N, 32), ceil_div(d.M, 128), nb * d.splitk);
    NT_CHECK(nt::launch_k(gemm_simt_kernel<128, 32, 16>, dim3(grid), dim3(256), 0, d));
  } else {
    dim3 grid(ceil_div(d.N, 64), ceil_div(d.M, 64), nb * d.splitk);
    NT_CHECK(nt::launch_k(gemm_simt_kernel<64, 64, 16>, dim3(grid), dim3(256), 0, d));
  }
  NT_LAUNCH_OK();
  return 0;
}

}  // namespace nt
