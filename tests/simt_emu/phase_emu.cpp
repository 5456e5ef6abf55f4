// CPU build of the phase-program kernel (numpy_transformer_b200/csrc/phase_kernel.cuh) under the SIMT emulator, plus the
// library's own recorder (phase_record.h).  TEST INFRASTRUCTURE: tests/test_phase_program_cpu.py drives it through ctypes
// with NumPy arrays standing in for device memory.  The emu_* entry points mirror the C-ABI calls the recorder intercepts.
#include "simt_emu.h"
#include "../../numpy_transformer_b200/csrc/phase_kernel.cuh"
#include "../../numpy_transformer_b200/csrc/phase_record.h"
#include "../../numpy_transformer_b200/csrc/gpt_block.h"

namespace emu { Launch* g_launch = nullptr; }

static ntph::Recorder g_rec;
static int g_err_flag = 0;

static void entry(const void* p) { ntph::run_program(*reinterpret_cast<const NtPhaseProgram*>(p)); }

extern "C" {

int emu_sizeof_phase() { return (int)sizeof(NtPhase); }
void emu_begin() { g_rec.clear(); g_err_flag = 0; }
int emu_phase_count() { return (int)g_rec.phases.size(); }
int emu_barrier_count() {
  int n = 0;
  for (const NtPhase& p : g_rec.phases) n += (p.flags & NT_PHF_SYNC_AFTER) ? 1 : 0;
  return n;
}
int emu_err_flag() { return g_err_flag; }
void emu_phase_info(int i, int* op, int* flags) { *op = g_rec.phases[i].op; *flags = (int)g_rec.phases[i].flags; }

// run what is recorded as launches of <= NT_PH_MAX_PHASES phases on `nctas` CTAs; all_barriers != 0 ignores the dependency
// tracking and puts a cluster barrier after every phase (the reference result for the tracked version); all_barriers == 2
// strips every cluster barrier (self-test of the emulator: the result must then be wrong)
int emu_run(int nctas, int reverse, int all_barriers) {
  size_t done = 0;
  while (done < g_rec.phases.size()) {
    static NtPhaseProgram prog;
    memset(&prog, 0, sizeof(prog));
    const size_t n = g_rec.phases.size() - done < NT_PH_MAX_PHASES ? g_rec.phases.size() - done : NT_PH_MAX_PHASES;
    prog.n = (int)n;
    prog.stamps = 0;
    for (size_t i = 0; i < n; i++) {
      prog.ph[i] = g_rec.phases[done + i];
      if (all_barriers == 1) prog.ph[i].flags |= NT_PHF_SYNC_AFTER;
      if (all_barriers == 2) prog.ph[i].flags &= ~NT_PHF_SYNC_AFTER;
    }
    const int rc = emu::run(entry, &prog, nctas, NT_PH_THREADS, NT_PH_SMEM_BYTES, reverse);
    if (rc) return rc;
    done += n;
  }
  g_rec.clear();
  return 0;
}

int emu_linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                   const float* residual) {
  return g_rec.linear_fwd(x, w, b, y, M, K, N, act, pre, residual, nullptr, nullptr) ? 0 : -1;
}
int emu_linear_ln_skinny(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* b, float* y, int M, int K,
                         int N, int act, const float* residual) {
  return g_rec.linear_fwd(x, w, b, y, M, K, N, act, nullptr, residual, ln_g, ln_b) ? 0 : -1;
}
int emu_linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                         const float* addend) {
  return g_rec.linear_bwd_input(dy, w, dx, M, K, N, act, pre, addend) ? 0 : -1;
}
int emu_linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
  return g_rec.linear_bwd_params(x, dy, dw, db, M, K, N) ? 0 : -1;
}
int emu_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M, int D,
                      float eps) {
  return g_rec.ln_fwd(x, gamma, beta, y, mean, rstd, M, D, eps, nullptr) ? 0 : -1;
}
int emu_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx,
                      float* dgamma, float* dbeta, int M, int D, const float* addend) {
  return g_rec.ln_bwd(dy, x, mean, rstd, gamma, dx, dgamma, dbeta, M, D, addend) ? 0 : -1;
}
int emu_attention_fwd(const float* qkv, int mask_kind, float* out, float* P, int B, int N, int H, int dh, const float* residual) {
  return g_rec.attn_fwd(qkv, nullptr, mask_kind, out, P, nullptr, B, N, H, dh, residual) ? 0 : -1;
}
int emu_attention_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, int B, int N, int H, int dh) {
  return g_rec.attn_bwd(dout, qkv, P, dqkv, B, N, H, dh) ? 0 : -1;
}
int emu_cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm, float* loss, float* row_loss,
                      float* grad, float* prob, int32_t* argmax, int rows, int cols) {
  return g_rec.cross_entropy(logits, labels, onehot, n_norm, loss, row_loss, grad, prob, argmax, rows, cols, &g_err_flag) ? 0 : -1;
}
int emu_embedding_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D, int vocab) {
  return g_rec.embed_fwd(idx, etok, epos, out, B, T, D, vocab, &g_err_flag) ? 0 : -1;
}
int emu_embedding_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D, int vocab) {
  return g_rec.embed_bwd(idx, dy, detok, depos, B, T, D, vocab, &g_err_flag) ? 0 : -1;
}
int emu_decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                     const int32_t* d_pos, int vocab) {
  return g_rec.decode_embed(idx, etok, epos, out, B, D, Tmax, d_pos, vocab, &g_err_flag) ? 0 : -1;
}
int emu_attention_decode_at(const float* qkv_new, float* kc, float* vc, float* out, int B, int Tmax, int H, int dh,
                            const int32_t* d_pos, int append) {
  return g_rec.attn_decode(qkv_new, kc, vc, out, B, Tmax, H, dh, d_pos, append) ? 0 : -1;
}
int emu_argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos) {
  return g_rec.argmax_next(logits, ld, cols, B, tok, hist, hist_ld, d_pos) ? 0 : -1;
}
int emu_increment(int32_t* counter) { return g_rec.increment(counter) ? 0 : -1; }

// gpt_block.h (what nt_gpt_block_fwd / nt_gpt_block_bwd issue) over the recorder
struct EmuOps {
  static int ln_fwd(const float* x, const float* g, const float* b, float* y, float* mean, float* rstd, int M, int D) {
    return g_rec.ln_fwd(x, g, b, y, mean, rstd, M, D, 1e-5f, nullptr) ? 0 : -1;
  }
  static int ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* g, float* dx, float* dg,
                    float* db, int M, int D, const float* addend) {
    return g_rec.ln_bwd(dy, x, mean, rstd, g, dx, dg, db, M, D, addend) ? 0 : -1;
  }
  static int linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                        const float* residual) {
    return g_rec.linear_fwd(x, w, b, y, M, K, N, act, pre, residual, nullptr, nullptr) ? 0 : -1;
  }
  static int linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                              const float* addend) {
    return g_rec.linear_bwd_input(dy, w, dx, M, K, N, act, pre, addend) ? 0 : -1;
  }
  static int linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
    return g_rec.linear_bwd_params(x, dy, dw, db, M, K, N) ? 0 : -1;
  }
  static int attn_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, int B, int N, int H, int dh) {
    return g_rec.attn_fwd(qkv, mask, mask_kind, out, P, nullptr, B, N, H, dh, nullptr) ? 0 : -1;
  }
  static int attn_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, float*, int B, int N, int H, int dh, int) {
    return g_rec.attn_bwd(dout, qkv, P, dqkv, B, N, H, dh) ? 0 : -1;
  }
};
int emu_gpt_block_fwd(const NtGptBlock* blk, const float* x, float* y, int B, int T, int D, int H, int mask_kind) {
  return ntgpt::block_fwd<EmuOps>(*blk, x, y, B, T, D, H, nullptr, mask_kind);
}
size_t emu_gpt_block_bwd_workspace(int B, int T, int D, int H) { return ntgpt::bwd_workspace_floats(B, T, D, H); }
int emu_gpt_block_bwd(const NtGptBlock* blk, const float* x, const float* dy, float* dx, float* ws, int B, int T, int D, int H,
                      int mask_kind) {
  return ntgpt::block_bwd<EmuOps>(*blk, x, dy, dx, ws, B, T, D, H, mask_kind);
}

}  // extern "C"
