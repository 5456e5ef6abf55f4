// gpt/attdecoderblock.py:31-111 as two calls (`nt_gpt_block_fwd / nt_gpt_block_bwd`, include/ntb200.h): the op sequence of
// the decoder block — MLP first, then masked multi-head attention — over an `Ops` policy.  In the library the policy
// forwards to the C-ABI entry points inside a fuse region (so a <= 16-row block is ONE phase-program launch each way and
// any other shape runs the same per-op kernels the layer objects use); in the CPU emulator build (tests/simt_emu) it
// forwards to the recorder, which is how the pointer plumbing below is tested without a GPU.
#pragma once
#include <cstddef>
#include "../../include/ntb200.h"

namespace ntgpt {

// floats of workspace nt_gpt_block_bwd needs: d0 [M,D] | dqkv [M,3D] | d1 [M,D] | qkvdelta [M,D] | d2 [M,4D] | d3 [M,D]
// | the attention backward's dS scratch [B,H,T,T] when T > 64
inline size_t bwd_workspace_floats(int B, int T, int D, int H) {
  const size_t M = (size_t)B * T;
  return M * D * 11 + (T > 64 ? (size_t)B * H * T * T : 0);
}

template <class Ops>
int block_fwd(const NtGptBlock& k, const float* x, float* y, int B, int T, int D, int H, const float* mask, int mask_kind) {
  const int M = B * T, F = 4 * D, dh = D / H;
  int rc;
  // h = x + FC1(ReLU(FC0(LN(x))))                                      attdecoderblock.py:37-45
  if ((rc = Ops::ln_fwd(x, k.ln_g, k.ln_b, k.out0, k.mean0, k.rstd0, M, D))) return rc;
  if ((rc = Ops::linear_fwd(k.out0, k.w0, k.b0, k.out2, M, D, F, NT_ACT_RELU, k.pre2, nullptr))) return rc;
  if ((rc = Ops::linear_fwd(k.out2, k.w1, k.b1, k.out6, M, F, D, NT_ACT_NONE, nullptr, x))) return rc;
  // out = h + FCout(MHA(LN1(h), mask))                                 attdecoderblock.py:47-75
  if ((rc = Ops::ln_fwd(k.out6, k.ln1_g, k.ln1_b, k.outnorm, k.mean1, k.rstd1, M, D))) return rc;
  if ((rc = Ops::linear_fwd(k.outnorm, k.wqkv, k.bqkv, k.qkv, M, D, 3 * D, NT_ACT_NONE, nullptr, nullptr))) return rc;
  if ((rc = Ops::attn_fwd(k.qkv, mask, mask_kind, k.rkk, k.P, B, T, H, dh))) return rc;
  return Ops::linear_fwd(k.rkk, k.wo, k.bo, y, M, D, D, NT_ACT_NONE, nullptr, k.out6);
}

template <class Ops>
int block_bwd(const NtGptBlock& k, const float* x, const float* dy, float* dx, float* ws, int B, int T, int D, int H, int mask_kind) {
  const int M = B * T, F = 4 * D, dh = D / H;
  const size_t MD = (size_t)M * D;
  float* d0 = ws;
  float* dqkv = d0 + MD;
  float* d1 = dqkv + 3 * MD;
  float* qkvdelta = d1 + MD;
  float* d2 = qkvdelta + MD;
  float* d3 = d2 + 4 * MD;
  float* scratch = T > 64 ? d3 + MD : nullptr;
  int rc;
  // fc_out, attention, qkv, norm1 (+ delta)                             attdecoderblock.py:79-105
  if ((rc = Ops::linear_bwd_params(k.rkk, dy, k.d_wo, k.d_bo, M, D, D))) return rc;
  if ((rc = Ops::linear_bwd_input(dy, k.wo, d0, M, D, D, NT_ACT_NONE, nullptr, nullptr))) return rc;
  if ((rc = Ops::attn_bwd(d0, k.qkv, k.P, dqkv, scratch, B, T, H, dh, mask_kind))) return rc;
  if ((rc = Ops::linear_bwd_params(k.outnorm, dqkv, k.d_wqkv, k.d_bqkv, M, D, 3 * D))) return rc;
  if ((rc = Ops::linear_bwd_input(dqkv, k.wqkv, d1, M, D, 3 * D, NT_ACT_NONE, nullptr, nullptr))) return rc;
  if ((rc = Ops::ln_bwd(d1, k.out6, k.mean1, k.rstd1, k.ln1_g, qkvdelta, k.d_ln1_g, k.d_ln1_b, M, D, dy))) return rc;
  // fc1 (+ ReLU'), fc0, norm (+ qkvdelta)                               attdecoderblock.py:106-111
  if ((rc = Ops::linear_bwd_params(k.out2, qkvdelta, k.d_w1, k.d_b1, M, F, D))) return rc;
  if ((rc = Ops::linear_bwd_input(qkvdelta, k.w1, d2, M, F, D, NT_ACT_RELU, k.pre2, nullptr))) return rc;
  if ((rc = Ops::linear_bwd_params(k.out0, d2, k.d_w0, k.d_b0, M, D, F))) return rc;
  if ((rc = Ops::linear_bwd_input(d2, k.w0, d3, M, D, F, NT_ACT_NONE, nullptr, nullptr))) return rc;
  return Ops::ln_bwd(d3, x, k.mean0, k.rstd0, k.ln_g, dx, k.d_ln_g, k.d_ln_b, M, D, qkvdelta);
}

}  // namespace ntgpt
