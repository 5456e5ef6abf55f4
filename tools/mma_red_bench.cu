// Development aid: issue rates that decide the design of the fused ViT-block kernels (vit_block.cu):
//   (1) mma.sync m16n8k8 TF32 and m16n8k16 BF16 throughput per SM at 4/8/16 warps,
//   (2) red.global.add(.v4).f32 throughput when ~100 CTAs accumulate into the same gradient buffer.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a tools/mma_red_bench.cu -o tools/mma_red_bench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

__global__ void mma_tf32_kernel(float* out, int iters) {
  float d[8][4];
  for (int i = 0; i < 8; i++) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void mma_bf16_kernel(float* out, int iters) {
  float d[8][4];
  for (int i = 0; i < 8; i++) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  uint32_t a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = a0 * 5;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++)
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(d[i][0]), "+f"(d[i][1]), "+f"(d[i][2]), "+f"(d[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0.f;
  for (int i = 0; i < 8; i++) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void ffma_kernel(float* out, int iters) {
  float d[16];
  for (int i = 0; i < 16; i++) d[i] = threadIdx.x * 0.001f + i;
  float a = 1.0001f, b = 0.0001f * threadIdx.x;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 16; i++) d[i] = fmaf(d[i], a, b);
  }
  float s = 0.f;
  for (int i = 0; i < 16; i++) s += d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// every CTA adds `n` floats into the same buffer, `reps` different buffers (one per "layer")
__global__ void red_v4_kernel(float* buf, int n, int reps) {
  for (int r = 0; r < reps; r++) {
    float* p = buf + (size_t)r * n;
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4)
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
  }
}
__global__ void red_v2_kernel(float* buf, int n, int reps) {
  for (int r = 0; r < reps; r++) {
    float* p = buf + (size_t)r * n;
    for (int i = threadIdx.x * 2; i < n; i += blockDim.x * 2)
      asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(p + i), "f"(1.f), "f"(2.f) : "memory");
  }
}
__global__ void red_s_kernel(float* buf, int n, int reps) {
  for (int r = 0; r < reps; r++) {
    float* p = buf + (size_t)r * n;
    for (int i = threadIdx.x; i < n; i += blockDim.x) atomicAdd(p + i, 1.f);
  }
}
// same but each CTA starts at a rotated offset so CTAs do not hit the same address at the same time
__global__ void red_v4_rot_kernel(float* buf, int n, int reps) {
  const int rot = (int)(((long long)blockIdx.x * n / gridDim.x) & ~3);
  for (int r = 0; r < reps; r++) {
    float* p = buf + (size_t)r * n;
    for (int i = threadIdx.x * 4; i < n; i += blockDim.x * 4) {
      int j = i + rot; if (j >= n) j -= n;
      asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p + j), "f"(1.f), "f"(2.f), "f"(3.f), "f"(4.f) : "memory");
    }
  }
}

template <typename F>
static float time_ms(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  float best = 1e9f;
  for (int i = 0; i < 3; i++) {
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount;
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("device %s, %d SMs, max clock %d MHz\n", p.name, sms, khz / 1000);
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 2);
  const int iters = 20000;
  for (int warps : {4, 8, 16}) {
    float ms = time_ms([&] { mma_tf32_kernel<<<sms, warps * 32>>>(out, iters); });
    double mma = (double)sms * warps * iters * 8;
    printf("mma.sync m16n8k8 tf32  %2d warps/SM: %.3f ms  %.1f TFLOP/s dense  (%.3f mma/ns/SM)\n", warps, ms, mma * 2048 / ms * 1e-9, mma / sms / ms * 1e-6);
    ms = time_ms([&] { mma_bf16_kernel<<<sms, warps * 32>>>(out, iters); });
    printf("mma.sync m16n8k16 bf16 %2d warps/SM: %.3f ms  %.1f TFLOP/s dense  (%.3f mma/ns/SM)\n", warps, ms, mma * 4096 / ms * 1e-9, mma / sms / ms * 1e-6);
    ms = time_ms([&] { ffma_kernel<<<sms, warps * 32>>>(out, iters); });
    printf("ffma                   %2d warps/SM: %.3f ms  %.1f TFLOP/s\n", warps, ms, (double)sms * warps * 32 * iters * 16 * 2 / ms * 1e-9);
  }
  const int n = 64512, reps = 6;     // one ViT block's weights (96x288 + 96x192 + 192x96), six blocks
  float* buf; cudaMalloc(&buf, sizeof(float) * (size_t)n * reps); cudaMemset(buf, 0, sizeof(float) * (size_t)n * reps);
  for (int grid : {25, 50, 100, 148}) {
    float a = time_ms([&] { red_v4_kernel<<<grid, 256>>>(buf, n, reps); });
    float b = time_ms([&] { red_v4_rot_kernel<<<grid, 256>>>(buf, n, reps); });
    float c = time_ms([&] { red_v2_kernel<<<grid, 256>>>(buf, n, reps); });
    float d = time_ms([&] { red_s_kernel<<<grid, 256>>>(buf, n, reps); });
    double el = (double)grid * n * reps;
    printf("red.add %3d CTAs x %d floats x %d: v4 %.1f us (%.1f Gfloat/s) | v4 rotated %.1f us (%.1f) | v2 %.1f us (%.1f) | scalar %.1f us (%.1f)\n",
           grid, n, reps, a * 1e3, el / a * 1e-6, b * 1e3, el / b * 1e-6, c * 1e3, el / c * 1e-6, d * 1e3, el / d * 1e-6);
  }
  printf("err=%d\n", (int)cudaGetLastError());
  return 0;
}
