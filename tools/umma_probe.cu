// Probe of the tcgen05 operand conventions the fused ViT kernels (csrc/vit_block_tc.cu) rely on, each checked against a
// CPU product of the same bf16 values:
//   1 K-major A x K-major B, operands written to shared memory by ordinary threads (not TMA), A rows 64..127 = whatever
//     follows in memory (their D rows are never read), K = 96 = 1.5 swizzle chunks
//   2 MN-major A x MN-major B from the same [token][feature] planes (weight-gradient form dW = X^T dY)
//   3 per-head score product: K-major A and B starting 64 bytes into the 128-byte swizzle row (head-padded dh = 32)
//   4 P V: K-major A, MN-major B with a 64-byte start offset and N = 32 (half a chunk)
//   5 P^T dO: MN-major A whose second 64-wide chunk is garbage, MN-major B with offset
//   6 N = 192 accumulator at a non-zero TMEM column, accumulation across two separately committed groups
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/umma_probe tools/umma_probe.cu
#include "../numpy_transformer_b200/csrc/umma.cuh"
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cstring>

using namespace nt::um;

struct Case {
  int N, steps, a_mn, b_mn, dcol;
  uint32_t a_start[16], b_start[16];      // byte offsets (into the A / B images) of each K=16 step
  uint32_t a_stride, b_stride;            // MN-major chunk strides
  uint32_t a_bytes, b_bytes;              // image sizes
  int split_commit;                       // commit after the first half of the steps, wait, then issue the rest
  int M;                                  // 0 -> 128
};

__global__ void __launch_bounds__(128) probe_kernel(const uint8_t* __restrict__ aimg, const uint8_t* __restrict__ bimg, Case c,
                                                    float* __restrict__ out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((c.a_bytes + 1023) & ~1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (uint32_t i = tid * 16; i < c.a_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sa + i) = *reinterpret_cast<const uint4*>(aimg + i);
  for (uint32_t i = tid * 16; i < c.b_bytes; i += 128 * 16) *reinterpret_cast<uint4*>(sb + i) = *reinterpret_cast<const uint4*>(bimg + i);
  fence_async_smem();                     // generic-proxy stores -> visible to the tensor core
  if (tid == 0) { mbar_init(&bar, 1); mbar_fence_init(); }
  if (warp == 0) tmem_alloc<512>(&tslot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  if (tid == 0) {
    const uint32_t idesc = idesc_bf16(c.M ? c.M : 128, c.N, c.a_mn, c.b_mn);
    uint32_t parity = 0;
    for (int s = 0; s < c.steps; s++) {
      const uint32_t aa = smem_u32(sa) + c.a_start[s], bb = smem_u32(sb) + c.b_start[s];
      const uint64_t ad = c.a_mn ? desc_mnmajor(aa, c.a_stride) : desc_kmajor(aa);
      const uint64_t bd = c.b_mn ? desc_mnmajor(bb, c.b_stride) : desc_kmajor(bb);
      mma_bf16(tbase + c.dcol, ad, bd, idesc, s > 0);
      if (c.split_commit && s == c.steps / 2 - 1) {
        mma_commit(&bar);
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
      }
    }
    mma_commit(&bar);
    mbar_wait(&bar, parity);
  }
  __syncthreads();
  tc_fence_after();
  for (int c0 = 0; c0 < c.N; c0 += 8) {
    uint32_t r[8];
    tmem_ld8(tmem_addr(tbase, warp * 32, c.dcol + c0), r);
    tmem_wait_ld();
    for (int j = 0; j < 8; j++) out[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

static uint16_t f2bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  u += 0x7FFF + ((u >> 16) & 1);
  return (uint16_t)(u >> 16);
}
static float bf2f(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct Plane {          // host image of a [rows][cols] matrix in the chunked swizzled layout
  int rows, cols;
  uint32_t stride;
  std::vector<uint8_t> img;
  std::vector<float> val;      // the bf16-rounded values, row-major [rows][cols]
  Plane(int r, int c, uint32_t st, size_t bytes, unsigned seed) : rows(r), cols(c), stride(st), img(bytes), val((size_t)r * c) {
    for (size_t i = 0; i + 1 < bytes; i += 2) { img[i] = 0xC0; img[i + 1] = 0x7F; }   // bf16 NaN everywhere first
    srand(seed);
    for (int i = 0; i < r; i++)
      for (int j = 0; j < c; j++) {
        const float f = (float)(rand() % 2001 - 1000) / 500.0f;
        const uint16_t h = f2bf(f);
        val[(size_t)i * c + j] = bf2f(h);
        memcpy(&img[plane_off(i, j, st)], &h, 2);
      }
  }
  float at(int i, int j) const { return val[(size_t)i * cols + j]; }
};

static int run_case(const char* name, const Case& c, const Plane& A, const Plane& B, int Mreal, int Nreal,
                    float (*ref)(const Plane&, const Plane&, int, int)) {
  uint8_t *da, *db;
  float* dout;
  cudaMalloc(&da, c.a_bytes); cudaMalloc(&db, c.b_bytes); cudaMalloc(&dout, 128 * c.N * 4);
  cudaMemcpy(da, A.img.data(), c.a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, B.img.data(), c.b_bytes, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0, 128 * c.N * 4);
  const size_t smem = ((c.a_bytes + 1023) & ~1023u) + c.b_bytes + 2048;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  probe_kernel<<<1, 128, smem>>>(da, db, c, dout);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("case %s: CUDA error %s\n", name, cudaGetErrorString(e)); return 1; }
  std::vector<float> out(128 * c.N);
  cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
  double worst = 0, scale = 0;
  int nan = 0;
  for (int i = 0; i < Mreal; i++)
    for (int j = 0; j < Nreal; j++) {
      const float r = ref(A, B, i, j), g = out[i * c.N + j];
      if (g != g) nan++;
      worst = fmax(worst, fabs((double)r - g));
      scale = fmax(scale, fabs((double)r));
    }
  const int ok = nan == 0 && worst <= 1e-3 * scale;
  printf("case %-28s %s  max|err| %.3e (max|ref| %.3e, NaNs %d)\n", name, ok ? "OK  " : "FAIL", worst, scale, nan);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return !ok;
}

int main() {
  int fails = 0;
  {  // 1: K-major x K-major, K = 96
    Plane A(64, 96, 8192, 2 * 8192, 1), B(96, 96, 96 * 128, 2 * 96 * 128, 2);
    Case c{}; c.N = 96; c.steps = 6; c.a_bytes = 2 * 8192 + 8192; c.b_bytes = 2 * 96 * 128;
    A.img.resize(c.a_bytes, 0);
    for (int s = 0; s < 6; s++) { c.a_start[s] = (s / 4) * 8192 + (s % 4) * 32; c.b_start[s] = (s / 4) * 96 * 128 + (s % 4) * 32; }
    fails += run_case("1 kmajor x kmajor K=96", c, A, B, 64, 96, [](const Plane& A, const Plane& B, int i, int j) {
      float s = 0; for (int k = 0; k < 96; k++) s += A.at(i, k) * B.at(j, k); return s; });
  }
  {  // 2: MN-major x MN-major (dW = X^T Y), X [64 tok][128 feat], Y [64 tok][96]
    Plane X(64, 128, 8192, 2 * 8192, 3), Y(64, 96, 8192, 2 * 8192, 4);
    Case c{}; c.N = 96; c.steps = 4; c.a_mn = c.b_mn = 1; c.a_stride = c.b_stride = 8192; c.a_bytes = c.b_bytes = 2 * 8192;
    for (int s = 0; s < 4; s++) c.a_start[s] = c.b_start[s] = s * 2048;
    fails += run_case("2 mnmajor x mnmajor (dW)", c, X, Y, 128, 96, [](const Plane& X, const Plane& Y, int i, int j) {
      float s = 0; for (int t = 0; t < 64; t++) s += X.at(t, i) * Y.at(t, j); return s; });
  }
  {  // 3: scores of head 1: Q, K planes [64][128], 64-byte start offset, K = 32
    Plane Q(64, 128, 8192, 2 * 8192, 5), K(64, 128, 8192, 2 * 8192, 6);
    Case c{}; c.N = 64; c.steps = 2; c.a_bytes = 3 * 8192; c.b_bytes = 2 * 8192;
    Q.img.resize(c.a_bytes, 0);
    for (int s = 0; s < 2; s++) c.a_start[s] = c.b_start[s] = 64 + s * 32;
    fails += run_case("3 head-offset scores", c, Q, K, 64, 64, [](const Plane& Q, const Plane& K, int i, int j) {
      float s = 0; for (int d = 0; d < 32; d++) s += Q.at(i, 32 + d) * K.at(j, 32 + d); return s; });
  }
  {  // 4: O = P V_h, head 3 (chunk 1, +64 bytes), N = 32
    Plane P(64, 64, 8192, 8192, 7), V(64, 128, 8192, 2 * 8192, 8);
    Case c{}; c.N = 32; c.steps = 4; c.b_mn = 1; c.b_stride = 8192; c.a_bytes = 2 * 8192; c.b_bytes = 2 * 8192;
    P.img.resize(c.a_bytes, 0);
    for (int s = 0; s < 4; s++) { c.a_start[s] = s * 32; c.b_start[s] = 8192 + 64 + s * 2048; }
    fails += run_case("4 P V (mnmajor B, offset)", c, P, V, 64, 32, [](const Plane& P, const Plane& V, int i, int j) {
      float s = 0; for (int t = 0; t < 64; t++) s += P.at(i, t) * V.at(t, 96 + j); return s; });
  }
  {  // 5: dV = P^T dO_h: MN-major A (second chunk garbage), MN-major B with offset, head 1
    Plane P(64, 64, 8192, 2 * 8192, 9), dO(64, 128, 8192, 2 * 8192, 10);
    Case c{}; c.N = 32; c.steps = 4; c.a_mn = c.b_mn = 1; c.a_stride = c.b_stride = 8192; c.a_bytes = c.b_bytes = 2 * 8192;
    for (int s = 0; s < 4; s++) { c.a_start[s] = s * 2048; c.b_start[s] = 64 + s * 2048; }
    fails += run_case("5 P^T dO (mnmajor A garbage)", c, P, dO, 64, 32, [](const Plane& P, const Plane& dO, int i, int j) {
      float s = 0; for (int q = 0; q < 64; q++) s += P.at(q, i) * dO.at(q, 32 + j); return s; });
  }
  {  // 6: N = 192 at TMEM column 288, two commit groups
    Plane A(64, 96, 8192, 2 * 8192, 11), B(192, 96, 192 * 128, 2 * 192 * 128, 12);
    Case c{}; c.N = 192; c.steps = 6; c.dcol = 288; c.split_commit = 1; c.a_bytes = 3 * 8192; c.b_bytes = 2 * 192 * 128;
    A.img.resize(c.a_bytes, 0);
    for (int s = 0; s < 6; s++) { c.a_start[s] = (s / 4) * 8192 + (s % 4) * 32; c.b_start[s] = (s / 4) * 192 * 128 + (s % 4) * 32; }
    fails += run_case("6 N=192 @col 288, 2 commits", c, A, B, 64, 192, [](const Plane& A, const Plane& B, int i, int j) {
      float s = 0; for (int k = 0; k < 96; k++) s += A.at(i, k) * B.at(j, k); return s; });
  }
  {  // 7: where do the rows of an M = 64 instruction land in TMEM?  (accumulator pre-filled by an M=128 product of zeros)
    Plane A(64, 64, 8192, 8192, 21), B(64, 64, 8192, 8192, 22);
    Case c{}; c.N = 64; c.steps = 4; c.M = 64; c.a_bytes = 2 * 8192; c.b_bytes = 8192;
    A.img.resize(c.a_bytes, 0);
    for (int s = 0; s < 4; s++) c.a_start[s] = c.b_start[s] = s * 32;
    uint8_t *da, *db; float* dout;
    cudaMalloc(&da, c.a_bytes); cudaMalloc(&db, c.b_bytes); cudaMalloc(&dout, 128 * c.N * 4);
    cudaMemcpy(da, A.img.data(), c.a_bytes, cudaMemcpyHostToDevice);
    cudaMemcpy(db, B.img.data(), c.b_bytes, cudaMemcpyHostToDevice);
    const size_t smem = 2 * 8192 + 8192 + 2048;
    probe_kernel<<<1, 128, smem>>>(da, db, c, dout);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) printf("case 7: CUDA error %s\n", cudaGetErrorString(e));
    else {
      std::vector<float> out(128 * 64);
      cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost);
      printf("case 7 M=64: TMEM lane <- A row (by matching the whole row of D): ");
      for (int lane = 0; lane < 128; lane++) {
        int found = -1;
        for (int i = 0; i < 64 && found < 0; i++) {
          bool ok = true;
          for (int j = 0; j < 64 && ok; j++) {
            float s = 0; for (int k = 0; k < 64; k++) s += A.at(i, k) * B.at(j, k);
            ok = fabsf(s - out[lane * 64 + j]) <= 1e-3f * (1.0f + fabsf(s));
          }
          if (ok) found = i;
        }
        if (found >= 0) printf("%d<-%d ", lane, found);
      }
      printf("\n");
    }
  }
  printf(fails ? "PROBE FAILED (%d)\n" : "PROBE OK\n", fails);
  return fails;
}
