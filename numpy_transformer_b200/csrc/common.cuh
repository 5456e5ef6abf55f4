// Internal helpers shared by the kernels of libntb200.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#include "../../include/ntb200.h"

namespace nt {

extern cudaStream_t g_stream;
extern int g_sm_count;
extern int g_gemm_mode;
extern long long g_launches;
extern bool g_capturing;
extern bool g_use_pdl;
// Device-side argument errors (token id >= vocabulary, label >= classes): kernels cannot return a status, so they set
// this word (pinned, host-mapped) and skip the access; nt_sync / nt_d2h report it at the next synchronisation.
extern int* g_dev_error;
enum { NT_DEVERR_TOKEN = 1, NT_DEVERR_LABEL = 2 };
void set_error(const char* fmt, ...);
bool ensure_init();

#define NT_CHECK(expr)                                                              \
  do {                                                                              \
    cudaError_t _e = (expr);                                                        \
    if (_e != cudaSuccess) {                                                        \
      nt::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                     \
    }                                                                               \
  } while (0)

#define NT_REQUIRE(cond, ...)                   \
  do {                                          \
    if (!(cond)) {                              \
      nt::set_error(__VA_ARGS__);               \
      return 2;                                 \
    }                                           \
  } while (0)

// every kernel launch goes through this so gpu_launches is counted honestly
#define NT_LAUNCH_OK()                                                 \
  do {                                                                 \
    if (!nt::g_capturing) nt::g_launches++;                            \
    cudaError_t _e = cudaGetLastError();                               \
    if (_e != cudaSuccess) {                                           \
      nt::set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
      return 1;                                                        \
    }                                                                  \
  } while (0)

#define NT_ENTRY() do { if (!nt::ensure_init()) return 3; } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// exact-erf GELU and its derivative (net/activation.py:139-148)
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float gelu_grad_f(float x) {
  float s = x * 0.70710678118654752f;
  return 0.5f + 0.5f * erff(s) + x * 0.3989422804014327f * expf(-s * s);
}

// Programmatic dependent launch: every kernel of this library starts with pdl_prologue(), so it can be
// launched while its predecessor in the stream is still draining (the ~2 us dependent-launch gap of a
// 130-kernel latency-bound step is then hidden) and only blocks where it first touches global memory.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = g_stream;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// epilogue description shared by the SIMT and tcgen05 GEMMs
struct Epilogue {
  const float* bias = nullptr;      // [N], added first
  int act = 0;                      // forward activation applied after bias (NT_ACT_*)
  float* pre_out = nullptr;         // store pre-activation (row stride ldc)
  int dact = 0;                     // multiply by act'(pre_in) (backward fusion)
  const float* pre_in = nullptr;    // [M,N] row stride ldc
  const float* addend = nullptr;    // [M,N] row stride ldc, added last
  float alpha = 1.0f;               // scale on the accumulator
  int atomic = 0;                   // 1: atomicAdd into C (split-K / accumulate semantics)
  float* colsum = nullptr;          // [N] += column sums of the MN-major B operand (db of the weight-gradient GEMM)
  void* out_hi = nullptr;           // optional: the final output also as bf16 hi / lo planes [M][ldc] (operand of the next
  void* out_lo = nullptr;           //           BF16x3 GEMM, so no separate split pass reads C back)
};

// C(m,n) = alpha * sum_k A(m,k) B(k,n), generic strides, two-level batch (z = z0*nb1 + z1)
struct GemmDesc {
  const float* A; const float* B; float* C;
  int M, N, K;
  long long sAm, sAk, sBk, sBn, ldc;
  int nb0 = 1, nb1 = 1;
  long long bA0 = 0, bA1 = 0, bB0 = 0, bB1 = 0, bC0 = 0, bC1 = 0;
  int splitk = 1;
  // optional pre-split operands (bf16 hi/lo, contraction dim contiguous): A_b[M][lda_b], B_b[N][ldb_b] -> BF16x3 tcgen05 path
  const void* A_hi = nullptr; const void* A_lo = nullptr; const void* B_hi = nullptr; const void* B_lo = nullptr;
  long long lda_b = 0, ldb_b = 0;
  bool a_b_mn = false, b_b_mn = false;   // split operand stored [K][M] / [K][N] (MN-major) instead of [M][K] / [N][K]
  bool* colsum_done = nullptr;       // set by the engine that fused ep.colsum
  Epilogue ep;
};
int gemm_simt(const GemmDesc& d);
int gemm_dispatch(const GemmDesc& d);   // picks tcgen05 when the mode and shape allow
int gemm_tc(const GemmDesc& d, int passes, bool* handled);

}  // namespace nt
