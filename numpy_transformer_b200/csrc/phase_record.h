// Host-side recorder of phase programs (see phase_prog.h): eligibility of an op, its argument slots, and the dependency
// tracking that decides where the program needs a cluster barrier.  Plain C++ without CUDA so that the CPU emulator
// build (tests/simt_emu) records programs with exactly the code the library uses.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>
#include "phase_prog.h"

namespace ntph {

struct Range {
  uintptr_t lo, hi;
};

class Recorder {
 public:
  std::vector<NtPhase> phases;

  void clear() {
    phases.clear();
    open_r_.clear();
    open_w_.clear();
  }
  bool empty() const { return phases.empty(); }

  static int row_tile(int M) { return M <= 1 ? 1 : (M <= 2 ? 2 : (M <= 4 ? 4 : (M <= 8 ? 8 : 16))); }

  // ---- one method per op; false = the program cannot express this call (the caller launches what is recorded and runs
  // the op's own kernel)
  bool linear_fwd(const float* x, const float* w, const float* b, float* y, int M, int K, int N, int act, float* pre,
                  const float* residual, const float* ln_g, const float* ln_b) {
    if (M < 1 || M > NT_PH_MAX_ROWS || K < 1 || N < 1 || act < 0 || act > 2) return false;
    const int MT = row_tile(M), XS = MT >= 4 ? MT + 4 : MT, RC = MT < 4 ? MT : 4;
    const size_t smem = sizeof(float) * ((((size_t)K * XS + 3) & ~(size_t)3) + (size_t)(NT_PH_THREADS / 32) * RC * 32 * 4);
    if (smem > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_LINEAR_FWD);
    ph.i[0] = M; ph.i[1] = K; ph.i[2] = N; ph.i[3] = act;
    set(ph, 0, x); set(ph, 1, w); set(ph, 2, b); set(ph, 3, y); set(ph, 4, pre); set(ph, 5, residual); set(ph, 6, ln_g); set(ph, 7, ln_b);
    begin();
    rd(x, (size_t)M * K); rd(w, (size_t)K * N); rd(b, N); rd(residual, (size_t)M * N); rd(ln_g, K); rd(ln_b, K);
    wr(y, (size_t)M * N); wr(pre, (size_t)M * N);
    return commit(ph);
  }

  bool linear_bwd_input(const float* dy, const float* w, float* dx, int M, int K, int N, int act, const float* pre,
                        const float* addend) {
    if (M < 1 || M > NT_PH_MAX_ROWS || K < 1 || N < 1 || act < 0 || act > 2 || (act && !pre)) return false;
    const int MT = row_tile(M), NP = (N + 3) & ~3;
    if (sizeof(float) * (size_t)MT * NP > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_LINEAR_BWD_INPUT);
    ph.i[0] = M; ph.i[1] = K; ph.i[2] = N; ph.i[3] = act;
    set(ph, 0, dy); set(ph, 1, w); set(ph, 2, dx); set(ph, 3, pre); set(ph, 4, addend);
    begin();
    rd(dy, (size_t)M * N); rd(w, (size_t)K * N); rd(act ? pre : nullptr, (size_t)M * K); rd(addend, (size_t)M * K);
    wr(dx, (size_t)M * K);
    return commit(ph);
  }

  bool linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N) {
    if (M < 1 || M > NT_PH_MAX_ROWS || K < 1 || N < 1) return false;
    const int MT = row_tile(M), NP = (N + 3) & ~3;
    if (sizeof(float) * ((size_t)MT * NP + (size_t)K * MT) > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_LINEAR_BWD_PARAMS);
    ph.i[0] = M; ph.i[1] = K; ph.i[2] = N;
    set(ph, 0, x); set(ph, 1, dy); set(ph, 2, dw); set(ph, 3, db);
    begin();
    rd(x, (size_t)M * K); rd(dy, (size_t)M * N);
    wr(dw, (size_t)K * N); wr(db, N);
    return commit(ph);
  }

  bool ln_fwd(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd, int M, int D,
              float eps, const float* post_add) {
    if (post_add || M < 1 || M > 64 || D < 1) return false;
    NtPhase ph = blank(NT_PH_LN_FWD);
    ph.i[0] = M; ph.i[1] = D; ph.f[0] = eps;
    set(ph, 0, x); set(ph, 1, gamma); set(ph, 2, beta); set(ph, 3, y); set(ph, 4, mean); set(ph, 5, rstd);
    begin();
    rd(x, (size_t)M * D); rd(gamma, D); rd(beta, D);
    wr(y, (size_t)M * D); wr(mean, M); wr(rstd, M);
    return commit(ph);
  }

  bool ln_bwd(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, float* dx,
              float* dgamma, float* dbeta, int M, int D, const float* addend) {
    if (M < 1 || M > 64 || D < 1) return false;
    NtPhase ph = blank(NT_PH_LN_BWD);
    ph.i[0] = M; ph.i[1] = D;
    set(ph, 0, dy); set(ph, 1, x); set(ph, 2, mean); set(ph, 3, rstd); set(ph, 4, gamma); set(ph, 5, dx); set(ph, 6, dgamma);
    set(ph, 7, dbeta); set(ph, 8, addend);
    begin();
    rd(dy, (size_t)M * D); rd(x, (size_t)M * D); rd(mean, M); rd(rstd, M); rd(gamma, D); rd(addend, (size_t)M * D);
    wr(dx, (size_t)M * D); wr(dgamma, D); wr(dbeta, D);
    return commit(ph);
  }

  bool attn_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S, int B, int N, int H, int dh,
                const float* residual) {
    if (mask || S || mask_kind < 0 || mask_kind > 1 || B < 1 || N < 1 || N > 32 || B * N > 64 || H < 1 || dh < 4 || (dh & 3) || dh > 128)
      return false;
    const size_t smem = sizeof(float) * (3 * (size_t)N * (dh + 4) + (size_t)N * (N + 1));
    if (smem > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_ATTN_FWD);
    ph.i[0] = B; ph.i[1] = N; ph.i[2] = H; ph.i[3] = dh; ph.i[4] = mask_kind;
    set(ph, 0, qkv); set(ph, 1, out); set(ph, 2, P); set(ph, 3, residual);
    const size_t D = (size_t)H * dh;
    begin();
    rd(qkv, (size_t)B * N * 3 * D); rd(residual, (size_t)B * N * D);
    wr(out, (size_t)B * N * D); wr(P, (size_t)B * H * N * N);
    return commit(ph);
  }

  bool attn_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, int B, int N, int H, int dh) {
    if (!P || B < 1 || N < 1 || N > 32 || B * N > 64 || H < 1 || dh < 4 || (dh & 3) || dh > 128) return false;
    const size_t smem = sizeof(float) * (4 * (size_t)N * (dh + 4) + 2 * (size_t)N * (N + 1));
    if (smem > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_ATTN_BWD);
    ph.i[0] = B; ph.i[1] = N; ph.i[2] = H; ph.i[3] = dh;
    set(ph, 0, dout); set(ph, 1, qkv); set(ph, 2, P); set(ph, 3, dqkv);
    const size_t D = (size_t)H * dh;
    begin();
    rd(dout, (size_t)B * N * D); rd(qkv, (size_t)B * N * 3 * D); rd(P, (size_t)B * H * N * N);
    wr(dqkv, (size_t)B * N * 3 * D);
    return commit(ph);
  }

  bool cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm, float* loss, float* row_loss,
                     float* grad, float* prob, int32_t* argmax, int rows, int cols, int* err) {
    if (rows < 1 || rows > 64 || cols < 1 || !row_loss || !loss || !(n_norm > 0.f)) return false;
    const float inv_n = 1.0f / n_norm;
    NtPhase ph = blank(NT_PH_CE_ROWS);
    ph.i[0] = rows; ph.i[1] = cols; ph.f[0] = inv_n;
    set(ph, 0, logits); set(ph, 1, labels); set(ph, 2, onehot); set(ph, 3, row_loss); set(ph, 4, grad); set(ph, 5, prob);
    set(ph, 6, argmax); set(ph, 7, err);
    begin();
    rd(logits, (size_t)rows * cols); rd(labels, rows); rd(onehot, (size_t)rows * cols);
    wr(row_loss, rows); wr(grad, (size_t)rows * cols); wr(prob, (size_t)rows * cols); wr(argmax, rows);
    if (!commit(ph)) return false;
    NtPhase pr = blank(NT_PH_CE_REDUCE);
    pr.i[0] = rows; pr.f[0] = inv_n;
    set(pr, 0, row_loss); set(pr, 1, loss);
    begin();
    rd(row_loss, rows);
    wr(loss, 1);
    return commit(pr);
  }

  bool embed_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D, int vocab, int* err) {
    if (B < 1 || T < 1 || B * T > 64 || D < 1 || vocab < 1) return false;
    NtPhase ph = blank(NT_PH_EMBED_FWD);
    ph.i[0] = B * T; ph.i[1] = T; ph.i[2] = D; ph.i[3] = vocab;
    set(ph, 0, idx); set(ph, 1, etok); set(ph, 2, epos); set(ph, 3, out); set(ph, 4, err);
    begin();
    rd(idx, (size_t)B * T); rd(etok, (size_t)vocab * D); rd(epos, (size_t)T * D);
    wr(out, (size_t)B * T * D);
    return commit(ph);
  }

  bool embed_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D, int vocab, int* err) {
    if (B < 1 || T < 1 || B * T > 64 || D < 1 || vocab < 1) return false;
    NtPhase ph = blank(NT_PH_EMBED_BWD);
    ph.i[0] = B; ph.i[1] = T; ph.i[2] = D; ph.i[3] = vocab;
    set(ph, 0, idx); set(ph, 1, dy); set(ph, 2, detok); set(ph, 3, depos); set(ph, 4, err);
    begin();
    rd(idx, (size_t)B * T); rd(dy, (size_t)B * T * D);
    wr(detok, (size_t)vocab * D); wr(depos, (size_t)T * D);
    return commit(ph);
  }

  bool decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                    const int32_t* d_pos, int vocab, int* err) {
    if (B < 1 || B > 64 || D < 1 || Tmax < 1 || vocab < 1 || !d_pos) return false;
    NtPhase ph = blank(NT_PH_DECODE_EMBED);
    ph.i[0] = B; ph.i[1] = D; ph.i[2] = Tmax; ph.i[3] = vocab;
    set(ph, 0, idx); set(ph, 1, etok); set(ph, 2, epos); set(ph, 3, out); set(ph, 4, d_pos); set(ph, 5, err);
    begin();
    rd(idx, B); rd(etok, (size_t)vocab * D); rd(epos, (size_t)Tmax * D); rd(d_pos, 1);
    wr(out, (size_t)B * D);
    return commit(ph);
  }

  bool attn_decode(const float* qkv_new, float* kc, float* vc, float* out, int B, int Tmax, int H, int dh, const int32_t* d_pos,
                   int append) {
    if (B < 1 || B > 64 || Tmax < 1 || H < 1 || dh < 1 || !d_pos) return false;
    const int NW = NT_PH_THREADS / 32;
    if (sizeof(float) * ((size_t)dh + (size_t)NW * dh + 2 * NW + Tmax) > NT_PH_SMEM_BYTES) return false;
    NtPhase ph = blank(NT_PH_ATTN_DECODE);
    ph.i[0] = B; ph.i[1] = Tmax; ph.i[2] = H; ph.i[3] = dh; ph.i[4] = append ? 1 : 0;
    set(ph, 0, qkv_new); set(ph, 1, kc); set(ph, 2, vc); set(ph, 3, out); set(ph, 4, d_pos);
    const size_t D = (size_t)H * dh;
    begin();
    rd(qkv_new, (size_t)B * 3 * D); rd(d_pos, 1);
    if (append) { wr(kc, (size_t)B * Tmax * D); wr(vc, (size_t)B * Tmax * D); }
    else { rd(kc, (size_t)B * Tmax * D); rd(vc, (size_t)B * Tmax * D); }
    wr(out, (size_t)B * D);
    return commit(ph);
  }

  bool argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos) {
    if (B < 1 || B > 64 || cols < 1 || ld < cols || !d_pos || !tok) return false;
    NtPhase ph = blank(NT_PH_ARGMAX_NEXT);
    ph.i[0] = ld; ph.i[1] = cols; ph.i[2] = B; ph.i[3] = hist_ld;
    set(ph, 0, logits); set(ph, 1, tok); set(ph, 2, hist); set(ph, 3, d_pos);
    begin();
    rd(logits, (size_t)(B - 1) * ld + cols); rd(d_pos, 1);
    wr(tok, B); wr(hist, (size_t)B * hist_ld);
    return commit(ph);
  }

  bool increment(int32_t* counter) {
    if (!counter) return false;
    NtPhase ph = blank(NT_PH_INCREMENT);
    set(ph, 0, counter);
    begin();
    wr(counter, 1);
    return commit(ph);
  }

 private:
  std::vector<Range> open_r_, open_w_;      // what the phases since the last barrier read / wrote
  std::vector<Range> cur_r_, cur_w_;

  static NtPhase blank(int op) {
    NtPhase ph;
    memset(&ph, 0, sizeof(ph));
    ph.op = op;
    return ph;
  }
  template <typename T>
  static void set(NtPhase& ph, int k, const T* p) { ph.p[k] = (uint64_t)reinterpret_cast<uintptr_t>(p); }
  void begin() {
    cur_r_.clear();
    cur_w_.clear();
  }
  template <typename T>
  void rd(const T* p, size_t n) {
    if (p && n) cur_r_.push_back(Range{reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + n * sizeof(T)});
  }
  template <typename T>
  void wr(T* p, size_t n) {
    if (p && n) cur_w_.push_back(Range{reinterpret_cast<uintptr_t>(p), reinterpret_cast<uintptr_t>(p) + n * sizeof(T)});
  }
  static bool overlap(const std::vector<Range>& a, const std::vector<Range>& b) {
    for (const Range& x : a)
      for (const Range& y : b)
        if (x.lo < y.hi && y.lo < x.hi) return true;
    return false;
  }
  // A phase that touches what an earlier phase since the last barrier wrote (RAW / WAW), or writes what one read (WAR),
  // needs the barrier in front of it: flag the previous phase.  Everything else shares a barrier interval.
  // An op whose output overlaps one of its own inputs (in-place use) is refused: the CTAs of a phase read and write
  // their slices concurrently.
  bool commit(const NtPhase& ph) {
    if (overlap(cur_w_, cur_r_)) return false;
    const bool dep = overlap(cur_r_, open_w_) || overlap(cur_w_, open_w_) || overlap(cur_w_, open_r_);
    if (dep && !phases.empty()) {
      phases.back().flags |= NT_PHF_SYNC_AFTER;
      open_r_.clear();
      open_w_.clear();
    }
    open_r_.insert(open_r_.end(), cur_r_.begin(), cur_r_.end());
    open_w_.insert(open_w_.end(), cur_w_.begin(), cur_w_.end());
    phases.push_back(ph);
    return true;
  }
};

}  // namespace ntph
