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

// Phase programs (phase_prog.cu): while g_fusing, the entry points of the ops a program can express hand their arguments
// to the matching fuse_* hook.  0 = recorded (nothing launched), NT_FUSE_PASS = not expressible: whatever was pending has
// been launched and the caller runs its own kernel, anything else = an error.  Every other kernel launch (launch_k) and
// stream operation first launches what is pending, so program order is stream order.
extern bool g_fusing;
enum { NT_FUSE_PASS = -1 };
int fuse_flush();
int fuse_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                    const float* residual, const float* ln_g, const float* ln_b);
int fuse_linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                          const float* addend);
int fuse_linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N);
int fuse_ln_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M, int D, float eps,
                const float* post_add);
int fuse_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma,
                float* dbeta, int M, int D, const float* addend);
int fuse_attn_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N, int H, int dh,
                  const float* residual);
int fuse_attn_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, int B, int N, int H, int dh);
int fuse_cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm, float* loss, float* row_loss,
                       float* grad, float* prob, int32_t* argmax, int rows, int cols);
int fuse_embed_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D, int vocab);
int fuse_embed_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D, int vocab);
int fuse_decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                      const int32_t* d_pos, int vocab);
int fuse_attn_decode(const float* qkv_new, float* kc, float* vc, float* out, int B, int Tmax, int H, int dh, const int32_t* d_pos,
                     int append);
int fuse_argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos);
int fuse_increment(int32_t* counter);

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

// in an entry point a phase program can express:  NT_FUSE_TRY(nt::fuse_xxx(args));
#define NT_FUSE_TRY(call)                          \
  do {                                             \
    if (nt::g_fusing) {                            \
      const int _r = (call);                       \
      if (_r != nt::NT_FUSE_PASS) return _r;       \
    }                                              \
  } while (0)
// in an entry point that enqueues stream work of its own (copies, memsets, collectives)
#define NT_FUSE_FLUSH()                            \
  do {                                             \
    if (nt::g_fusing) {                            \
      const int _r = nt::fuse_flush();             \
      if (_r) return _r;                           \
    }                                              \
  } while (0)

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
  if (g_fusing && fuse_flush() != 0) return cudaErrorLaunchFailure;     // pending phases go first (text in nt_last_error)
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
