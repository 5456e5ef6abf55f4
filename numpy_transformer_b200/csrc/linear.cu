// Linear layer entry points (net/fullconnect.py) -> GEMM descriptors.
#include "common.cuh"
#include <cstdlib>

namespace nt {

// db[n] += sum_m dy[m,n]   (net/fullconnect.py:122,125)
__global__ void colsum_kernel(const float* __restrict__ dy, float* __restrict__ db, int M, int N, int rows_per_block) {
  pdl_prologue();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  if (c < N)
    for (int r = r0 + threadIdx.y; r < r1; r += 8) s += __ldg(dy + (long long)r * N + c);
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) t += red[i][threadIdx.x];
    atomicAdd(db + c, t);
  }
}

// ---- skinny forward: M <= 8 rows (one decode step of the KV-cache generator, tiny heads).  A 128-row tensor-core tile
// would be >= 94 % padding and walk K serially in one CTA (20-50 us measured for 1 x 384 x {384..2350}); here the weight
// panel is streamed once by N/32 CTAs, 32 k-groups x 8 lanes x float4 each, exact fp32 FMAs, and the bias / activation /
// residual epilogue of net/fullconnect.py:96-104 follows the in-CTA reduction.
// ln_g / ln_b (may be NULL): the rows are LayerNorm-ed first (net/layernorm.py:100-114, biased variance, eps 1e-5) — every
// CTA normalises its own copy of the <= 8 rows, which is cheaper than a separate launch in the decode step.
template <int MR, int CL, bool VEC>      // CL lanes x float4 = 4*CL columns per CTA, 256/CL k-groups
__global__ void __launch_bounds__(256) linear_skinny_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ b, float* __restrict__ y, int M, int K,
                                                            int N, int act, float* __restrict__ pre,
                                                            const float* __restrict__ residual,
                                                            const float* __restrict__ ln_g, const float* __restrict__ ln_b) {
  pdl_prologue();
  constexpr int COLS = 4 * CL, KG = 256 / CL;
  extern __shared__ __align__(16) float sk[];
  float* sx = sk;                          // [MR][K]
  float* red = sk + ((MR * K + 3) & ~3);   // [8 warps][MR][COLS]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, cl = tid % CL, kg = tid / CL;
  const int n0 = blockIdx.x * COLS + cl * 4;
  for (int e = tid; e < MR * K; e += 256) sx[e] = e < M * K ? x[e] : 0.f;
  __syncthreads();
  if (ln_g) {
    for (int r = warp; r < M; r += 8) {      // one warp per row: two-pass mean / variance like layernorm.cu
      float* row = sx + r * K;
      float s = 0.f;
      for (int k = lane; k < K; k += 32) s += row[k];
      const float mean = warp_sum(s) / K;
      float q = 0.f;
      for (int k = lane; k < K; k += 32) { const float d = row[k] - mean; q += d * d; }
      const float rstd = rsqrtf(warp_sum(q) / K + 1e-5f);
      for (int k = lane; k < K; k += 32) row[k] = (row[k] - mean) * rstd * ln_g[k] + ln_b[k];
    }
    __syncthreads();
  }
  float acc[MR][4];
#pragma unroll
  for (int r = 0; r < MR; r++) acc[r][0] = acc[r][1] = acc[r][2] = acc[r][3] = 0.f;
  if (n0 < N) {
#pragma unroll 4
    for (int k = kg; k < K; k += KG) {
      float wv[4];
      const float* wr = w + (long long)k * N + n0;
      if (VEC) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(wr));
        wv[0] = t.x; wv[1] = t.y; wv[2] = t.z; wv[3] = t.w;
      } else {
#pragma unroll
        for (int c = 0; c < 4; c++) wv[c] = n0 + c < N ? __ldg(wr + c) : 0.f;
      }
#pragma unroll
      for (int r = 0; r < MR; r++) {
        const float xv = sx[r * K + k];
#pragma unroll
        for (int c = 0; c < 4; c++) acc[r][c] = fmaf(xv, wv[c], acc[r][c]);
      }
    }
  }
  // k-groups of one warp sit CL lanes apart: butterfly over the upper lane bits, then one partial per warp in shared memory
#pragma unroll
  for (int off = 16; off >= CL; off >>= 1)
#pragma unroll
    for (int r = 0; r < MR; r++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[r][c] += __shfl_xor_sync(0xffffffffu, acc[r][c], off);
  if (lane < CL) {
#pragma unroll
    for (int r = 0; r < MR; r++)
      *reinterpret_cast<float4*>(red + (warp * MR + r) * COLS + lane * 4) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
  }
  __syncthreads();
  for (int e = tid; e < MR * COLS; e += 256) {
    const int r = e / COLS, c = e % COLS, n = blockIdx.x * COLS + c;
    if (r >= M || n >= N) continue;
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 8; g++) v += red[(g * MR + r) * COLS + c];
    if (b) v += b[n];
    const long long o = (long long)r * N + n;
    if (pre) pre[o] = v;
    if (act == NT_ACT_RELU) v = fmaxf(v, 0.f);
    else if (act == NT_ACT_GELU) v = gelu_f(v);
    if (residual) v += residual[o];
    y[o] = v;
  }
}

template <int MR, int CL>
static int linear_skinny_launch_cl(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act,
                                   float* pre, const float* residual, const float* ln_g, const float* ln_b) {
  const size_t smem = sizeof(float) * ((((size_t)MR * K + 3) & ~(size_t)3) + 8 * MR * 4 * CL);
  const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  auto kern = vec ? linear_skinny_kernel<MR, CL, true> : linear_skinny_kernel<MR, CL, false>;
  static size_t configured[2] = {0, 0};
  if (smem > 48 * 1024 && smem > configured[vec]) {
    NT_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured[vec] = smem;
  }
  NT_CHECK(launch_k(kern, dim3(ceil_div(N, 4 * CL)), dim3(256), smem, x, w, b, y, M, K, N, act, pre, residual, ln_g, ln_b));
  NT_LAUNCH_OK();
  return 0;
}

// Columns per CTA shrink with N so that the weight panel is pulled by ~50-100 SMs instead of N/32 (12 for N = 384):
// the decode-step GEMVs are latency-bound, not bandwidth-bound (1 x 1536 x 384: 8 us with 12 CTAs).
template <int MR>
static int linear_skinny_launch(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act,
                                float* pre, const float* residual, const float* ln_g = nullptr, const float* ln_b = nullptr) {
  if (N <= 768) return linear_skinny_launch_cl<MR, 2>(x, w, b, y, M, K, N, act, pre, residual, ln_g, ln_b);
  if (N <= 1536) return linear_skinny_launch_cl<MR, 4>(x, w, b, y, M, K, N, act, pre, residual, ln_g, ln_b);
  return linear_skinny_launch_cl<MR, 8>(x, w, b, y, M, K, N, act, pre, residual, ln_g, ln_b);
}

// ---- backward of the same <= 16-row Linears (the gpt_character configuration: 10 token rows).  Through the tensor-core
// tile these cost 18 us (dX) and 10 us (dW) per layer, almost all of it the fixed latency of a 128-row tile pipeline.
// dX[M][K] = dY[M][N] W[K][N]^T (x act'(pre), + addend): one warp per weight row k — the row is read once, coalesced, and
// dotted with the <= 16 rows of dY that sit in shared memory; exact fp32.
template <int MR>
__global__ void __launch_bounds__(256) linear_skinny_bwd_input_kernel(const float* __restrict__ dy, const float* __restrict__ w,
                                                                      float* __restrict__ dx, int M, int K, int N, int act,
                                                                      const float* __restrict__ pre, const float* __restrict__ addend) {
  pdl_prologue();
  extern __shared__ __align__(16) float sdy[];           // [MR][N]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int e = tid; e < MR * N; e += 256) sdy[e] = e < M * N ? dy[e] : 0.f;
  __syncthreads();
  const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(w) & 15) == 0);
  for (int k = blockIdx.x * 8 + warp; k < K; k += gridDim.x * 8) {
    const float* wr = w + (long long)k * N;
    float acc[MR];
#pragma unroll
    for (int r = 0; r < MR; r++) acc[r] = 0.f;
    if (vec) {
      for (int n = 4 * lane; n < N; n += 128) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(wr + n));
#pragma unroll
        for (int r = 0; r < MR; r++) {
          const float4 d = *reinterpret_cast<const float4*>(sdy + r * N + n);
          acc[r] = fmaf(t.x, d.x, fmaf(t.y, d.y, fmaf(t.z, d.z, fmaf(t.w, d.w, acc[r]))));
        }
      }
    } else {
      for (int n = lane; n < N; n += 32) {
        const float t = __ldg(wr + n);
#pragma unroll
        for (int r = 0; r < MR; r++) acc[r] = fmaf(t, sdy[r * N + n], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < MR; r++) acc[r] = warp_sum(acc[r]);
#pragma unroll
    for (int r = 0; r < MR; r++)
      if (lane == r && r < M) {
        const long long o = (long long)r * K + k;
        float v = acc[r];
        if (act == NT_ACT_RELU) v = pre[o] > 0.f ? v : 0.f;
        else if (act == NT_ACT_GELU) v *= gelu_grad_f(pre[o]);
        if (addend) v += addend[o];
        dx[o] = v;
      }
  }
}

// dW[K][N] += x[M][K]^T dY[M][N], db[N] += colsum(dY): a thread owns four columns of a strip of weight rows; dY lives in its
// registers, x in shared memory; plain read-modify-write (every element has one owner), which keeps the reference's
// accumulate-until-setzero semantics (fullconnect.py:118-122).
template <int MR>
__global__ void __launch_bounds__(128) linear_skinny_bwd_params_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                       float* __restrict__ dw, float* __restrict__ db, int M, int K,
                                                                       int N, int krows) {
  pdl_prologue();
  extern __shared__ __align__(16) float sx[];            // [MR][krows]
  const int k0 = blockIdx.y * krows, kn = min(krows, K - k0);
  for (int e = threadIdx.x; e < MR * krows; e += 128) {
    const int r = e / krows, kk = e % krows;
    sx[e] = (r < M && kk < kn) ? x[(long long)r * K + k0 + kk] : 0.f;
  }
  __syncthreads();
  const int n = 4 * (blockIdx.x * 128 + threadIdx.x);
  if (n >= N) return;
  const bool vec = (N % 4 == 0) && ((reinterpret_cast<uintptr_t>(dw) & 15) == 0) && ((reinterpret_cast<uintptr_t>(dy) & 15) == 0);
  float d[MR][4];
#pragma unroll
  for (int r = 0; r < MR; r++) {
#pragma unroll
    for (int c = 0; c < 4; c++) d[r][c] = (r < M && n + c < N) ? dy[(long long)r * N + n + c] : 0.f;
  }
  if (db && blockIdx.y == 0) {
#pragma unroll
    for (int c = 0; c < 4; c++)
      if (n + c < N) {
        float sum = 0.f;
#pragma unroll
        for (int r = 0; r < MR; r++) sum += d[r][c];
        db[n + c] += sum;
      }
  }
  for (int kk = 0; kk < kn; kk++) {
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int r = 0; r < MR; r++) {
      const float xv = sx[r * krows + kk];
      a0 = fmaf(xv, d[r][0], a0); a1 = fmaf(xv, d[r][1], a1); a2 = fmaf(xv, d[r][2], a2); a3 = fmaf(xv, d[r][3], a3);
    }
    float* o = dw + (long long)(k0 + kk) * N + n;
    if (vec) {
      float4 t = *reinterpret_cast<float4*>(o);
      t.x += a0; t.y += a1; t.z += a2; t.w += a3;
      *reinterpret_cast<float4*>(o) = t;
    } else {
      if (n < N) o[0] += a0;
      if (n + 1 < N) o[1] += a1;
      if (n + 2 < N) o[2] += a2;
      if (n + 3 < N) o[3] += a3;
    }
  }
}

template <int MR>
static int linear_skinny_bwd_input_launch(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                                          const float* addend) {
  const size_t smem = sizeof(float) * (size_t)MR * N;
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    NT_CHECK(cudaFuncSetAttribute(linear_skinny_bwd_input_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = smem;
  }
  const int grid = ceil_div(K, 8) < 2 * g_sm_count ? ceil_div(K, 8) : 2 * g_sm_count;
  NT_CHECK(launch_k(linear_skinny_bwd_input_kernel<MR>, dim3(grid), dim3(256), smem, dy, w, dx, M, K, N, act, pre, addend));
  NT_LAUNCH_OK();
  return 0;
}
template <int MR>
static int linear_skinny_bwd_params_launch(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
  const int krows = 4;                 // K / 4 x N / 512 CTAs: these launches are latency-bound, spread them over the SMs
  NT_CHECK(launch_k(linear_skinny_bwd_params_kernel<MR>, dim3(ceil_div(N, 512), ceil_div(K, krows)), dim3(128),
                    sizeof(float) * (size_t)MR * krows, x, dy, dw, db, M, K, N, krows));
  NT_LAUNCH_OK();
  return 0;
}

// M <= 16 (a decode step, the heads of tiny batches, the 10-row Linears of the gpt_character config) and the row panel
// fits shared memory
static bool linear_skinny_ok(int M, int K) {
  if (const char* e = getenv("NTB200_SKINNY")) if (e[0] == '0') return false;
  const int mr = M <= 1 ? 1 : (M <= 2 ? 2 : (M <= 4 ? 4 : (M <= 8 ? 8 : 16)));
  return M >= 1 && M <= 16 && sizeof(float) * ((size_t)mr * K + 4 + 32 * mr * 32) <= 200 * 1024;
}

int gemm_dispatch(const GemmDesc& d) {
  if (g_gemm_mode != 0) {
    bool handled = false;
    int rc = gemm_tc(d, g_gemm_mode == 1 ? 3 : 1, &handled);
    if (rc) return rc;
    if (handled) return 0;
  }
  return gemm_simt(d);
}

}  // namespace nt
using namespace nt;

extern "C" {

int nt_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act,
                  float* pre, const float* residual) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0, "nt_linear_fwd: bad shape M=%d K=%d N=%d", M, K, N);
  NT_REQUIRE(act >= 0 && act <= 2, "nt_linear_fwd: act %d", act);
  NT_FUSE_TRY(nt::fuse_linear_fwd(x, w, b, y, M, K, N, act, pre, residual, nullptr, nullptr));
  if (linear_skinny_ok(M, K)) {
    if (M <= 1) return linear_skinny_launch<1>(x, w, b, y, M, K, N, act, pre, residual);
    if (M <= 2) return linear_skinny_launch<2>(x, w, b, y, M, K, N, act, pre, residual);
    if (M <= 4) return linear_skinny_launch<4>(x, w, b, y, M, K, N, act, pre, residual);
    if (M <= 8) return linear_skinny_launch<8>(x, w, b, y, M, K, N, act, pre, residual);
    return linear_skinny_launch<16>(x, w, b, y, M, K, N, act, pre, residual);
  }
  GemmDesc d{};
  d.A = x; d.B = w; d.C = y;
  d.M = M; d.N = N; d.K = K;
  d.sAm = K; d.sAk = 1; d.sBk = N; d.sBn = 1; d.ldc = N;
  // Few output tiles but a long contraction (the ViT head: 100 x 4704 -> 96): one CTA per tile would walk
  // ~150 k-blocks serially.  Split K across CTAs with red.add into a zeroed buffer, then apply the
  // activation / residual in a second, tiny pass.
  const int tiles = ceil_div(M, 128) * ceil_div(N, 32);
  if (g_gemm_mode != 0 && tiles * 4 < g_sm_count && K >= 1024) {
    float* acc = (act != 0 && pre) ? pre : y;
    NT_CHECK(cudaMemsetAsync(acc, 0, sizeof(float) * (size_t)M * N, g_stream));
    d.C = acc;
    d.ep.bias = b;
    d.ep.atomic = 1;
    d.splitk = 2;                       // > 1: gemm_tc sizes the split from the SM count
    int rc = gemm_dispatch(d);
    if (rc) return rc;
    if (act == 0 && !residual) return 0;
    if (act != 0 && !pre) {             // activation in place on y
      rc = act == NT_ACT_RELU ? nt_relu_fwd(y, y, (size_t)M * N) : nt_gelu_fwd(y, y, (size_t)M * N);
    } else if (act != 0) {
      rc = act == NT_ACT_RELU ? nt_relu_fwd(pre, y, (size_t)M * N) : nt_gelu_fwd(pre, y, (size_t)M * N);
    }
    if (rc) return rc;
    if (residual) return nt_add(y, residual, y, (size_t)M * N);
    return 0;
  }
  d.ep.bias = b; d.ep.act = act; d.ep.pre_out = pre; d.ep.addend = residual;
  return gemm_dispatch(d);
}

int nt_linear_ln_skinny(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* b, float* y, int M,
                        int K, int N, int act, const float* residual) {
  NT_ENTRY();
  NT_REQUIRE(M >= 1 && M <= 16 && K > 0 && N > 0 && ln_g && ln_b, "nt_linear_ln_skinny: needs 1..16 rows and the LayerNorm parameters (M=%d)", M);
  NT_REQUIRE(act >= 0 && act <= 2, "nt_linear_ln_skinny: act %d", act);
  NT_FUSE_TRY(nt::fuse_linear_fwd(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b));
  NT_REQUIRE(linear_skinny_ok(M, K), "nt_linear_ln_skinny: K=%d does not fit the row panel", K);
  if (M <= 1) return linear_skinny_launch<1>(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b);
  if (M <= 2) return linear_skinny_launch<2>(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b);
  if (M <= 4) return linear_skinny_launch<4>(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b);
  if (M <= 8) return linear_skinny_launch<8>(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b);
  return linear_skinny_launch<16>(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b);
}

int nt_linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act,
                        const float* pre, const float* addend) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0, "nt_linear_bwd_input: bad shape M=%d K=%d N=%d", M, K, N);
  NT_REQUIRE(act == 0 || pre != nullptr, "nt_linear_bwd_input: act' needs the saved pre-activation");
  NT_FUSE_TRY(nt::fuse_linear_bwd_input(dy, w, dx, M, K, N, act, pre, addend));
  if (M >= 1 && linear_skinny_ok(M, N)) {
    if (M <= 4) return linear_skinny_bwd_input_launch<4>(dy, w, dx, M, K, N, act, pre, addend);
    if (M <= 8) return linear_skinny_bwd_input_launch<8>(dy, w, dx, M, K, N, act, pre, addend);
    return linear_skinny_bwd_input_launch<16>(dy, w, dx, M, K, N, act, pre, addend);
  }
  GemmDesc d{};
  // dx[M,K] = dy[M,N] . w[K,N]^T : contraction over N
  d.A = dy; d.B = w; d.C = dx;
  d.M = M; d.N = K; d.K = N;
  d.sAm = N; d.sAk = 1; d.sBk = 1; d.sBn = N; d.ldc = K;
  d.ep.dact = act; d.ep.pre_in = pre; d.ep.addend = addend;
  return gemm_dispatch(d);
}

int nt_linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0, "nt_linear_bwd_params: bad shape M=%d K=%d N=%d", M, K, N);
  if (M == 0) return 0;
  NT_FUSE_TRY(nt::fuse_linear_bwd_params(x, dy, dw, db, M, K, N));
  if (linear_skinny_ok(M, K)) {
    if (M <= 4) return linear_skinny_bwd_params_launch<4>(x, dy, dw, db, M, K, N);
    if (M <= 8) return linear_skinny_bwd_params_launch<8>(x, dy, dw, db, M, K, N);
    return linear_skinny_bwd_params_launch<16>(x, dy, dw, db, M, K, N);
  }
  GemmDesc d{};
  // dw[K,N] += x[M,K]^T . dy[M,N] : contraction over M (long), split across CTAs, atomic accumulate
  d.A = x; d.B = dy; d.C = dw;
  d.M = K; d.N = N; d.K = M;
  d.sAm = 1; d.sAk = K; d.sBk = N; d.sBn = 1; d.ldc = N;
  d.ep.atomic = 1;
  int tiles = ceil_div(K, 64) * ceil_div(N, 64);
  int want = ceil_div(2 * g_sm_count, tiles);
  int maxsplit = ceil_div(M, 128);
  d.splitk = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
  bool db_fused = false;
  static int fuse_db = -1;
  if (fuse_db < 0) { const char* e = getenv("NTB200_FUSE_DB"); fuse_db = e ? atoi(e) : 1; }
  d.ep.colsum = fuse_db ? db : nullptr;
  d.colsum_done = &db_fused;
  int rc = gemm_dispatch(d);
  if (rc) return rc;
  if (db && !db_fused) {
    int rows_per_block = 256;
    dim3 grid(ceil_div(N, 32), ceil_div(M, rows_per_block));
    NT_CHECK(nt::launch_k(colsum_kernel, dim3(grid), dim3(dim3(32, 8)), 0, dy, db, M, N, rows_per_block));
    NT_LAUNCH_OK();
  }
  return 0;
}

// ---- BF16x3 variants: operands pre-split by nt_split_bf16 (hi/lo bf16, contraction dim contiguous)
int nt_linear_fwd_bf16x3(const void* x_hi, const void* x_lo, const void* wt_hi, const void* wt_lo, const float* b, float* y,
                         int M, int K, int N, int act, float* pre, const float* residual, void* y_hi, void* y_lo) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0 && (K & 7) == 0, "nt_linear_fwd_bf16x3: bad shape M=%d K=%d N=%d (K %% 8)", M, K, N);
  NT_REQUIRE(act >= 0 && act <= 2, "nt_linear_fwd_bf16x3: act %d", act);
  if (M == 0) return 0;
  GemmDesc d{};
  d.C = y; d.M = M; d.N = N; d.K = K; d.ldc = N;
  d.A_hi = x_hi; d.A_lo = x_lo; d.lda_b = K;          // x  [M][K]
  d.B_hi = wt_hi; d.B_lo = wt_lo; d.ldb_b = K;        // w^T [N][K]
  d.ep.bias = b; d.ep.act = act; d.ep.pre_out = pre; d.ep.addend = residual;
  NT_REQUIRE((y_hi != nullptr) == (y_lo != nullptr), "nt_linear_fwd_bf16x3: pass both split output planes or neither");
  d.ep.out_hi = y_hi; d.ep.out_lo = y_lo;
  bool handled = false;
  int rc = gemm_tc(d, 3, &handled);
  if (rc) return rc;
  NT_REQUIRE(handled, "nt_linear_fwd_bf16x3: shape/epilogue not supported by the tensor-core engine (M=%d K=%d N=%d)", M, K, N);
  return 0;
}

int nt_linear_bwd_input_bf16x3(const void* dy_hi, const void* dy_lo, const void* w_hi, const void* w_lo, float* dx, int M, int K,
                               int N, int act, const float* pre, const float* addend, void* dx_hi, void* dx_lo) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0 && (N & 7) == 0, "nt_linear_bwd_input_bf16x3: bad shape M=%d K=%d N=%d (N %% 8)", M, K, N);
  if (M == 0) return 0;
  GemmDesc d{};
  d.C = dx; d.M = M; d.N = K; d.K = N; d.ldc = K;
  d.A_hi = dy_hi; d.A_lo = dy_lo; d.lda_b = N;        // dy [M][N]
  d.B_hi = w_hi; d.B_lo = w_lo; d.ldb_b = N;          // w  [K][N]: row k_in is output column k_in, contraction over N
  d.ep.dact = act; d.ep.pre_in = pre; d.ep.addend = addend;
  NT_REQUIRE((dx_hi != nullptr) == (dx_lo != nullptr), "nt_linear_bwd_input_bf16x3: pass both split output planes or neither");
  d.ep.out_hi = dx_hi; d.ep.out_lo = dx_lo;
  bool handled = false;
  int rc = gemm_tc(d, 3, &handled);
  if (rc) return rc;
  NT_REQUIRE(handled, "nt_linear_bwd_input_bf16x3: shape/epilogue not supported by the tensor-core engine (M=%d K=%d N=%d)", M, K, N);
  return 0;
}

int nt_linear_bwd_params_bf16x3(const void* xt_hi, const void* xt_lo, const void* dyt_hi, const void* dyt_lo, int ld_t,
                                const float* dy, float* dw, float* db, int M, int K, int N) {
  NT_ENTRY();
  NT_REQUIRE(M >= 0 && K > 0 && N > 0, "nt_linear_bwd_params_bf16x3: bad shape M=%d K=%d N=%d", M, K, N);
  if (M == 0) return 0;
  GemmDesc d{};
  d.C = dw; d.M = K; d.N = N; d.K = M; d.ldc = N;
  if (ld_t == 0) {
    // untransposed splits x [M][K], dy [M][N]: both operands are MN-major for this GEMM (contraction over rows)
    NT_REQUIRE((K & 7) == 0 && (N & 7) == 0, "nt_linear_bwd_params_bf16x3: K %d and N %d must be multiples of 8", K, N);
    d.A_hi = xt_hi; d.A_lo = xt_lo; d.lda_b = K; d.a_b_mn = true;
    d.B_hi = dyt_hi; d.B_lo = dyt_lo; d.ldb_b = N; d.b_b_mn = true;
  } else {
    NT_REQUIRE((ld_t & 7) == 0 && ld_t >= M, "nt_linear_bwd_params_bf16x3: transposed pitch %d (M=%d)", ld_t, M);
    d.A_hi = xt_hi; d.A_lo = xt_lo; d.lda_b = ld_t;     // x^T  [K][M]
    d.B_hi = dyt_hi; d.B_lo = dyt_lo; d.ldb_b = ld_t;   // dy^T [N][M]
  }
  d.ep.atomic = 1;                                    // dw += (accumulate-until-setzero, fullconnect.py:120-121)
  d.splitk = 2;
  bool handled = false;
  int rc = gemm_tc(d, 3, &handled);
  if (rc) return rc;
  NT_REQUIRE(handled, "nt_linear_bwd_params_bf16x3: shape not supported by the tensor-core engine (M=%d K=%d N=%d)", M, K, N);
  if (db) {
    const int rpb = 512;
    NT_CHECK(nt::launch_k(colsum_kernel, dim3(ceil_div(N, 32), ceil_div(M, rpb)), dim3(32, 8), 0, dy, db, M, N, rpb));
    NT_LAUNCH_OK();
  }
  return 0;
}

}  // extern "C"
