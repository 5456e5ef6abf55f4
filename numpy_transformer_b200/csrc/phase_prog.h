// Phase programs: the launch-bound GPT shapes (<= 16 token rows: the gpt_character step, one KV-cache decode step) as
// ONE kernel.  Between nt_fuse_begin / nt_fuse_end the C-ABI entry points of the per-op path do not launch; each call
// appends one NtPhase to a host-side program.  nt_fuse_end (or any op the program cannot express) launches
// `phase_program_kernel`: a single thread-block cluster that walks the phases in order, every CTA taking a slice of each
// phase's columns / weight rows / (sample, head) items, with a hardware cluster barrier (~0.2 us) where a later phase
// reads what an earlier one wrote instead of a kernel boundary (~4 us).  This file is shared by the host recorder
// (phase_prog.cu), the device code (phase_kernel.cuh) and the CPU SIMT emulator build of the same device code
// (tests/simt_emu), so it uses only plain C types.
#pragma once
#include <stdint.h>

enum NtPhaseOp {
  NT_PH_NOP = 0,
  NT_PH_LINEAR_FWD,         // nt_linear_fwd / nt_linear_ln_skinny      (net/fullconnect.py:99-106, net/layernorm.py:100-114)
  NT_PH_LINEAR_BWD_INPUT,   // nt_linear_bwd_input                      (net/fullconnect.py:114)
  NT_PH_LINEAR_BWD_PARAMS,  // nt_linear_bwd_params                     (net/fullconnect.py:118-125)
  NT_PH_LN_FWD,             // nt_layernorm_fwd                         (net/layernorm.py:88-114)
  NT_PH_LN_BWD,             // nt_layernorm_bwd                         (net/layernorm.py:116-135)
  NT_PH_ATTN_FWD,           // nt_attention_fwd, N <= 32                (gpt/attdecoderblock.py:52-68)
  NT_PH_ATTN_BWD,           // nt_attention_bwd, N <= 32                (gpt/attdecoderblock.py:79-99)
  NT_PH_CE_ROWS,            // nt_cross_entropy, per-row half           (net/loss.py:46-51)
  NT_PH_CE_REDUCE,          // nt_cross_entropy, the scalar
  NT_PH_EMBED_FWD,          // nt_embedding_fwd                         (PatchEmbed.py:132-138, net/Embedding.py:46)
  NT_PH_EMBED_BWD,          // nt_embedding_bwd                         (PatchEmbed.py:140-146)
  NT_PH_DECODE_EMBED,       // nt_decode_embed
  NT_PH_ATTN_DECODE,        // nt_attention_decode_at
  NT_PH_ARGMAX_NEXT,        // nt_argmax_next
  NT_PH_INCREMENT,          // nt_increment
  NT_PH_COUNT
};

#define NT_PHF_SYNC_AFTER 1u      // a later phase depends on this one (or on one before it): cluster barrier after it

#define NT_PH_MAX_PHASES 64       // per launch; a longer program is cut into several launches
#define NT_PH_THREADS 512         // threads per CTA of the program kernel
#define NT_PH_MAX_ROWS 16         // token rows a Linear phase takes
#define NT_PH_SMEM_BYTES (200 * 1024)

// 128 bytes; the program travels as a kernel parameter (no device-side descriptor buffer whose lifetime a captured
// graph would have to pin)
struct NtPhase {
  int32_t op;
  uint32_t flags;
  int32_t i[8];
  float f[2];
  uint64_t p[10];
};

struct NtPhaseProgram {
  int32_t n;
  int32_t pad;
  uint64_t stamps;      // debug (nt_fuse_debug_stamps): device int64 [phase][CTA][3] = SM clock at phase start / body end / barrier end
  NtPhase ph[NT_PH_MAX_PHASES];
};

// Per-op argument slots --------------------------------------------------------------------------------------------------
// LINEAR_FWD        p: x w b y pre residual ln_g ln_b            i: M K N act
// LINEAR_BWD_INPUT  p: dy w dx pre addend                        i: M K N act
// LINEAR_BWD_PARAMS p: x dy dw db                                i: M K N
// LN_FWD            p: x gamma beta y mean rstd                  i: M D              f: eps
// LN_BWD            p: dy x mean rstd gamma dx dgamma dbeta addend   i: M D
// ATTN_FWD          p: qkv out P residual                        i: B N H dh mask_kind
// ATTN_BWD          p: dout qkv P dqkv                           i: B N H dh
// CE_ROWS           p: logits labels onehot row_loss grad prob argmax err   i: rows cols    f: inv_n
// CE_REDUCE         p: row_loss loss                             i: rows             f: inv_n
// EMBED_FWD         p: idx etok epos out err                     i: BT T D vocab
// EMBED_BWD         p: idx dy detok depos err                    i: B T D vocab
// DECODE_EMBED      p: idx etok epos out d_pos err               i: B D Tmax vocab
// ATTN_DECODE       p: qkv_new kc vc out d_pos                   i: B Tmax H dh append
// ARGMAX_NEXT       p: logits tok hist d_pos                     i: ld cols B hist_ld
// INCREMENT         p: counter
