// Phase programs on the device (see phase_prog.h): between nt_fuse_begin and nt_fuse_end the per-op entry points record
// phases instead of launching; the program runs as ONE thread-block cluster (`phase_program_kernel`).  This is the fused
// decoder block SURVEY.md 8(b) asks for at the launch-bound GPT shapes — gpt/attdecoderblock.py:31-111 forward and
// backward, the stem / head / loss around it and a whole KV-cache decode step — expressed as a list of the same ops the
// layer objects already issue, so the Python layers stay what they are.
#include "common.cuh"
#include "phase_kernel.cuh"
#include "phase_record.h"
#include "gpt_block.h"
#include <cstdlib>

namespace nt {

bool g_fusing = false;
static int g_fuse_depth = 0;
static ntph::Recorder g_rec;
static int g_cluster = 0;                     // CTAs per cluster (= grid size); 0 until configured
static long long g_fuse_launches = 0, g_fuse_phases = 0, g_fuse_barriers = 0;
static void* g_stamps = nullptr;              // debug: per-phase SM clock stamps of the launches that follow

__global__ void __launch_bounds__(NT_PH_THREADS, 1) phase_program_kernel(const __grid_constant__ NtPhaseProgram prog) {
  ph_prologue();
  ntph::run_program(prog);
}

// Largest cluster the device can co-schedule with one CTA (200 KB of shared memory, 512 threads) per SM: 16 is the
// non-portable maximum on sm_100 (a GPC has 16-20 SMs), 8 the portable one.  NTB200_FUSE_CLUSTER overrides.
static int configure() {
  if (g_cluster) return 0;
  NT_CHECK(cudaFuncSetAttribute(phase_program_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, NT_PH_SMEM_BYTES));
  int want = 16;
  if (const char* e = getenv("NTB200_FUSE_CLUSTER")) {
    const int v = atoi(e);
    if (v >= 1 && v <= 16) want = v;
  }
  if (want > 8 && cudaFuncSetAttribute(phase_program_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
    cudaGetLastError();
    want = 8;
  }
  int pick = 0;
  for (int c = want; c >= 1; c >>= 1) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(c);
    cfg.blockDim = dim3(NT_PH_THREADS);
    cfg.dynamicSmemBytes = NT_PH_SMEM_BYTES;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, phase_program_kernel, &cfg) == cudaSuccess && n >= 1) { pick = c; break; }
    cudaGetLastError();
  }
  NT_REQUIRE(pick >= 1, "phase programs: no cluster size fits this device");
  g_cluster = pick;
  return 0;
}

// launch what is recorded (<= NT_PH_MAX_PHASES phases per launch); the recorder is empty afterwards whatever happens
int fuse_flush() {
  if (g_rec.empty()) return 0;
  std::vector<NtPhase> phases;
  phases.swap(g_rec.phases);
  g_rec.clear();
  if (configure()) return 1;
  static NtPhaseProgram prog;
  size_t done = 0;
  while (done < phases.size()) {
    const size_t n = phases.size() - done < (size_t)NT_PH_MAX_PHASES ? phases.size() - done : (size_t)NT_PH_MAX_PHASES;
    memset(&prog, 0, sizeof(prog));
    prog.n = (int)n;
    prog.stamps = (uint64_t)reinterpret_cast<uintptr_t>(g_stamps);
    for (size_t i = 0; i < n; i++) {
      prog.ph[i] = phases[done + i];
      if (prog.ph[i].flags & NT_PHF_SYNC_AFTER) g_fuse_barriers++;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(g_cluster);
    cfg.blockDim = dim3(NT_PH_THREADS);
    cfg.dynamicSmemBytes = NT_PH_SMEM_BYTES;
    cfg.stream = g_stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = g_cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t le = cudaLaunchKernelEx(&cfg, phase_program_kernel, prog);
    while (le != cudaSuccess && g_cluster > 1 && !g_capturing) {
      // the occupancy query granted a cluster the launch could not place (partitioned GPCs ...): settle for half
      cudaGetLastError();
      g_cluster >>= 1;
      cfg.gridDim = dim3(g_cluster);
      at[0].val.clusterDim.x = g_cluster;
      le = cudaLaunchKernelEx(&cfg, phase_program_kernel, prog);
    }
    NT_CHECK(le);
    NT_LAUNCH_OK();
    g_fuse_launches++;
    g_fuse_phases += (long long)n;
    done += n;
  }
  return 0;
}

// shared tail of the fuse_* hooks: recorded -> 0; not expressible -> launch what is pending, then let the caller run its
// own kernel (NT_FUSE_PASS) or report the launch failure
static inline int recorded_or_pass(bool ok) {
  if (ok) return 0;
  const int rc = fuse_flush();
  return rc ? rc : NT_FUSE_PASS;
}

int fuse_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                    const float* residual, const float* ln_g, const float* ln_b) {
  return recorded_or_pass(g_rec.linear_fwd(x, w, b, y, M, K, N, act, pre, residual, ln_g, ln_b));
}
int fuse_linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                          const float* addend) {
  return recorded_or_pass(g_rec.linear_bwd_input(dy, w, dx, M, K, N, act, pre, addend));
}
int fuse_linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
  return recorded_or_pass(g_rec.linear_bwd_params(x, dy, dw, db, M, K, N));
}
int fuse_ln_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M, int D, float eps,
                const float* post_add) {
  return recorded_or_pass(g_rec.ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, post_add));
}
int fuse_ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx, float* dgamma,
                float* dbeta, int M, int D, const float* addend) {
  return recorded_or_pass(g_rec.ln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend));
}
int fuse_attn_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N, int H, int dh,
                  const float* residual) {
  return recorded_or_pass(g_rec.attn_fwd(qkv, mask, mask_kind, out, P, S, B, N, H, dh, residual));
}
int fuse_attn_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, int B, int N, int H, int dh) {
  return recorded_or_pass(g_rec.attn_bwd(dout, qkv, P, dqkv, B, N, H, dh));
}
int fuse_cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm, float* loss, float* row_loss,
                       float* grad, float* prob, int32_t* argmax, int rows, int cols) {
  return recorded_or_pass(g_rec.cross_entropy(logits, labels, onehot, n_norm, loss, row_loss, grad, prob, argmax, rows, cols, g_dev_error));
}
int fuse_embed_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D, int vocab) {
  return recorded_or_pass(g_rec.embed_fwd(idx, etok, epos, out, B, T, D, vocab, g_dev_error));
}
int fuse_embed_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D, int vocab) {
  return recorded_or_pass(g_rec.embed_bwd(idx, dy, detok, depos, B, T, D, vocab, g_dev_error));
}
int fuse_decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                      const int32_t* d_pos, int vocab) {
  return recorded_or_pass(g_rec.decode_embed(idx, etok, epos, out, B, D, Tmax, d_pos, vocab, g_dev_error));
}
int fuse_attn_decode(const float* qkv_new, float* kc, float* vc, float* out, int B, int Tmax, int H, int dh, const int32_t* d_pos,
                     int append) {
  return recorded_or_pass(g_rec.attn_decode(qkv_new, kc, vc, out, B, Tmax, H, dh, d_pos, append));
}
int fuse_argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos) {
  return recorded_or_pass(g_rec.argmax_next(logits, ld, cols, B, tok, hist, hist_ld, d_pos));
}
int fuse_increment(int32_t* counter) { return recorded_or_pass(g_rec.increment(counter)); }

// the decoder block's op sequence (gpt_block.h) over the library's own entry points
struct LibOps {
  static int ln_fwd(const float* x, const float* g, const float* b, float* y, float* mean, float* rstd, int M, int D) {
    return nt_layernorm_fwd(x, g, b, y, mean, rstd, M, D, 1e-5f, nullptr, 0);          // net/layernorm.py:86
  }
  static int ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* g, float* dx, float* dg,
                    float* db, int M, int D, const float* addend) {
    return nt_layernorm_bwd(dy, x, mean, rstd, g, dx, dg, db, M, D, addend);
  }
  static int linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                        const float* residual) {
    return nt_linear_fwd(x, w, b, y, M, K, N, act, pre, residual);
  }
  static int linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                              const float* addend) {
    return nt_linear_bwd_input(dy, w, dx, M, K, N, act, pre, addend);
  }
  static int linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
    return nt_linear_bwd_params(x, dy, dw, db, M, K, N);
  }
  static int attn_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, int B, int N, int H, int dh) {
    return nt_attention_fwd(qkv, mask, mask_kind, out, P, nullptr, B, N, H, dh, nullptr);
  }
  static int attn_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, float* scratch, int B, int N, int H, int dh,
                      int mask_kind) {
    return nt_attention_bwd(dout, qkv, P, dqkv, scratch, B, N, H, dh, mask_kind);
  }
};

}  // namespace nt
using namespace nt;

extern "C" {

int nt_gpt_block_fwd(const NtGptBlock* blk, const float* x, float* y, int B, int T, int D, int H, const float* mask, int mask_kind) {
  NT_ENTRY();
  NT_REQUIRE(blk && x && y && B >= 0 && T > 0 && D > 0 && H > 0 && D % H == 0, "nt_gpt_block_fwd: bad arguments (D %d, heads %d)", D, H);
  NT_REQUIRE(mask_kind != NT_MASK_DENSE || mask, "nt_gpt_block_fwd: dense mask_kind needs a mask");
  if (B == 0) return 0;
  int rc = nt_fuse_begin();
  if (rc) return rc;
  rc = ntgpt::block_fwd<LibOps>(*blk, x, y, B, T, D, H, mask_kind == NT_MASK_DENSE ? mask : nullptr, mask_kind);
  const int rc_end = nt_fuse_end();
  return rc ? rc : rc_end;
}

int nt_gpt_block_bwd_workspace(int B, int T, int D, int H, size_t* floats) {
  NT_REQUIRE(floats && B >= 0 && T > 0 && D > 0 && H > 0, "nt_gpt_block_bwd_workspace: bad arguments");
  *floats = ntgpt::bwd_workspace_floats(B, T, D, H);
  return 0;
}

int nt_gpt_block_bwd(const NtGptBlock* blk, const float* x, const float* dy, float* dx, float* workspace, int B, int T, int D,
                     int H, int mask_kind) {
  NT_ENTRY();
  NT_REQUIRE(blk && x && dy && dx && workspace && B >= 0 && T > 0 && D > 0 && H > 0 && D % H == 0,
             "nt_gpt_block_bwd: bad arguments (D %d, heads %d)", D, H);
  if (B == 0) return 0;
  int rc = nt_fuse_begin();
  if (rc) return rc;
  rc = ntgpt::block_bwd<LibOps>(*blk, x, dy, dx, workspace, B, T, D, H, mask_kind);
  const int rc_end = nt_fuse_end();
  return rc ? rc : rc_end;
}

int nt_fuse_begin(void) {
  NT_ENTRY();
  if (g_fuse_depth++ == 0) {
    g_rec.clear();
    g_fusing = true;
  }
  return 0;
}

int nt_fuse_end(void) {
  NT_ENTRY();
  NT_REQUIRE(g_fuse_depth > 0, "nt_fuse_end without nt_fuse_begin");
  if (--g_fuse_depth > 0) return 0;
  const int rc = fuse_flush();
  g_fusing = false;
  return rc;
}

int nt_fuse_debug_stamps(void* dev_int64_buf) {
  g_stamps = dev_int64_buf;
  return 0;
}

int nt_fuse_stats(long long* launches, long long* phases, long long* barriers, int* cluster_size) {
  if (launches) *launches = g_fuse_launches;
  if (phases) *phases = g_fuse_phases;
  if (barriers) *barriers = g_fuse_barriers;
  if (cluster_size) *cluster_size = g_cluster;
  return 0;
}

}  // extern "C"
