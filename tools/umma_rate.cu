// How long does one tcgen05.mma (kind::f16, bf16 operands, cta_group::1, M = 128, K = 16) take as a function of N and of
// the operand layouts?  One CTA per SM issues a chain of CHAIN instructions on zeroed shared-memory operands (no loads, no
// epilogue) and measures issue -> commit with clock64.  Development aid behind DESIGN.md 4.6 / 4.1 ("an M = 128 MMA pays a
// fixed price for its A strip"): build with
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I numpy_transformer_b200/csrc tools/umma_rate.cu -o tools/umma_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma.cuh"

using namespace nt::um;

struct Cfg { int N, a_mn, b_mn, chain, kadv; };     // kadv: advance the operands by one K step per instruction (as a GEMM does)

__global__ void __launch_bounds__(128) rate_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  for (uint32_t i = threadIdx.x * 16; i < 160 * 1024; i += blockDim.x * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&tslot);
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (threadIdx.x == 0) {
    const uint32_t sa = smem_u32(smem), sbb = sa + 64 * 1024;     // A: 64 KB region, B: 96 KB region
    const uint32_t idesc = idesc_bf16(128, c.N, c.a_mn, c.b_mn);
    const uint64_t a0 = c.a_mn ? desc_mnmajor(sa, 8192) : desc_kmajor(sa);
    const uint64_t b0 = c.b_mn ? desc_mnmajor(sbb, 8192) : desc_kmajor(sbb);
    const uint32_t astep = (c.a_mn ? MNMAJ_STEP : KMAJ_STEP) >> 4, bstep = (c.b_mn ? MNMAJ_STEP : KMAJ_STEP) >> 4;
    uint32_t ph = 0;
    for (int rep = 0; rep < 3; rep++) {               // rep 0 warms up
      const long long t0 = clock64();
      for (int i = 0; i < c.chain; i++) {
        const int ks = c.kadv ? (i & 3) : 0;
        mma_bf16(tbase, a0 + (uint64_t)(ks * astep), b0 + (uint64_t)(ks * bstep), idesc, i ? 1u : 0u);
      }
      const long long t1 = clock64();
      mma_commit(&bar);
      mbar_wait(&bar, ph);
      ph ^= 1;
      const long long t2 = clock64();
      if (blockIdx.x == 0 && rep == 2) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tbase);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int Ns[] = {32, 64, 96, 128, 192, 256};
  printf("tcgen05.mma kind::f16 M=128 K=16, chain of 64 on every SM (%d CTAs): cycles per instruction (issue loop | until commit)\n", sms);
  for (int lay = 0; lay < 4; lay++) {
    const int a_mn = lay >> 1, b_mn = lay & 1;
    for (int kadv = 0; kadv < 2; kadv++) {
      printf("A %s, B %s, %s:", a_mn ? "MN-major" : "K-major ", b_mn ? "MN-major" : "K-major ", kadv ? "K steps 0..3 " : "same operands");
      for (int N : Ns) {
        Cfg c{N, a_mn, b_mn, 64, kadv};
        rate_kernel<<<sms, 128, 160 * 1024>>>(c, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf(" N=%d: %s\n", N, cudaGetErrorString(e)); return 1; }
        long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("  N=%d: %.0f | %.0f", N, h[0] / 64.0, h[1] / 64.0);
      }
      printf("\n");
    }
  }
  printf("peak at 8192 dense bf16 FLOP/clk/SM: N=64 32, N=128 64, N=256 128 cycles per instruction\n");
  return 0;
}
