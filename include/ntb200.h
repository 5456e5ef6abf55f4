/* ntb200.h — C ABI of libntb200.so, the B200 (sm_100a) backend for the training step of
 * ZouJiu1/numpy_transformer.
 *
 * The reference has no FFI: its "plugin API" is the duck-typed Python layer protocol
 * forward/backward/update/setzero/save_model/restore_model.  Each entry point
 * below is what the host-side mirror of one reference method binds through ctypes; the
 * reference method it replaces is cited as file:line (paths relative to the reference root).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; nt_last_error() gives the text.
 *   - all tensor arguments are DEVICE pointers to dense row-major fp32 unless stated otherwise;
 *     index tensors are int32.  Host pointers appear only in nt_h2d / nt_d2h.
 *   - work is enqueued on the library's single compute stream; nothing synchronises except
 *     nt_d2h, nt_sync and nt_event_elapsed_ms.  All ops are CUDA-graph capturable.
 *   - "accumulates" means += into the destination (the reference's *_delta buffers, which
 *     live until setzero()).
 */
#ifndef NTB200_H
#define NTB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- runtime */
int nt_init(int device);                 /* select device, create the compute stream (idempotent) */
int nt_shutdown(void);
const char* nt_last_error(void);
int nt_device_count(int* n);
int nt_sm_count(int* n);
int nt_malloc(void** dptr, size_t bytes);   /* stream-ordered size-class pool over cudaMalloc */
int nt_free(void* dptr);
int nt_pool_bytes(size_t* reserved);
int nt_host_alloc(void** hptr, size_t bytes);   /* pinned host memory for the e2e path */
int nt_host_free(void* hptr);
int nt_h2d(void* dst, const void* src, size_t bytes);
int nt_d2h(void* dst, const void* src, size_t bytes);     /* synchronises the stream */
int nt_d2h_async(void* dst, const void* src, size_t bytes);
int nt_d2d(void* dst, const void* src, size_t bytes);
int nt_memset(void* dst, int value, size_t bytes);
int nt_sync(void);
int nt_stream_handle(void** cuda_stream);
int nt_event_create(void** ev);
int nt_event_record(void* ev);
int nt_event_record_external(void* ev);                   /* capturable: becomes a graph event-record node */
int nt_event_elapsed_ms(void* ev0, void* ev1, float* ms);  /* synchronises on ev1 */
int nt_event_synchronize(void* ev);                        /* host waits for the event (pinned staging reuse) */
int nt_event_destroy(void* ev);
/* CUDA-graph capture of whatever is enqueued between begin and end (one step of the trainer
 * loop transformer_of_image.py:141-159 / gpt/gpt_train_potry3000.py:256-288). */
int nt_graph_begin(void);
int nt_graph_end(void** graph_exec, int* n_kernel_nodes);
int nt_graph_launch(void* graph_exec);
/* Fork/join of a side branch inside a captured step (no-ops outside capture): ops enqueued between
 * begin/end run concurrently with the main chain; join before anything consumes their results. */
int nt_set_side_branch(int enabled);      /* 0: capture everything on the main stream (clean per-op timing) */
int nt_side_begin(void);
int nt_side_end(void);
int nt_side_join(void);
int nt_graph_destroy(void* graph_exec);
/* number of this library's kernels launched so far (graph replays count their kernel nodes) */
int nt_launch_count(long long* n);
/* GEMM engine: 0 = fp32 SIMT (exact fp32), 1 = tcgen05 TF32 x3 split (fp32-class accuracy),
 * 2 = tcgen05 TF32 single pass.  Shapes the tensor-core path cannot take fall back to 0. */
int nt_set_gemm_mode(int mode);
int nt_get_gemm_mode(int* mode);
int nt_flush_l2(void);                   /* writes a 256 MB scratch buffer (bench hygiene) */

/* ---------------------------------------------------------------- Linear  (net/fullconnect.py) */
#define NT_ACT_NONE 0
#define NT_ACT_RELU 1
#define NT_ACT_GELU 2
/* fclayer.forward, net/fullconnect.py:99-106:  y[M,N] = act(x[M,K] @ w[K,N] + b[N]) + residual[M,N].
 * b, pre, residual may be NULL.  If pre != NULL the pre-activation is stored there
 * (GELU.forward net/activation.py:139-142 and ReLU.forward :11-14 fused as epilogues). */
int nt_linear_fwd(const float* x, const float* w, const float* b, float* y,
                  int M, int K, int N, int act, float* pre, const float* residual);
/* fclayer.backward input half, net/fullconnect.py:114:  dx[M,K] = (dy[M,N] @ w[K,N]^T) * act'(pre[M,K]) + addend[M,K].
 * act' is GELU.backward (net/activation.py:144-148) or ReLU.backward (:16-17) of the layer that
 * PRODUCED x; pre / addend may be NULL. */
/* Decode-step form (1..16 rows): y = act(LayerNorm(x; ln_g, ln_b) W + b) + residual in one launch (layernorm.py:100-114 followed
 * by fullconnect.py:99-106); nothing is saved for a backward. */
int nt_linear_ln_skinny(const float* x, const float* ln_g, const float* ln_b, const float* w, const float* b, float* y, int M,
                        int K, int N, int act, const float* residual);
int nt_linear_bwd_input(const float* dy, const float* w, float* dx,
                        int M, int K, int N, int act, const float* pre, const float* addend);
/* fclayer.backward parameter half, net/fullconnect.py:118-125:  dw[K,N] += x^T @ dy ; db[N] += colsum(dy).
 * db may be NULL. Accumulates. */
int nt_linear_bwd_params(const float* x, const float* dy, float* dw, float* db, int M, int K, int N);

/* BF16x3 engine for the large Linear shapes (fp32-class accuracy at twice the tensor rate of the TF32 split):
 * every fp32 operand is first split in HBM into hi = bf16(x), lo = bf16(x - hi) (optionally transposed so that the
 * contraction dimension is contiguous), then the GEMM accumulates lo*hi + hi*lo + hi*hi with tcgen05 kind::f16.
 * nt_split_bf16: x [rows, cols] fp32 -> planes hi / lo of bf16; transpose = 0: [rows][ld_out] (ld_out >= cols),
 * transpose = 1: [cols][ld_out] (ld_out >= rows, pad to a multiple of 8). */
int nt_split_bf16(const float* x, void* hi, void* lo, int rows, int cols, int ld_out, int transpose);
/* both layouts in one pass over x: hi / lo [rows][cols] and hi_t / lo_t [cols][ld_t] (a Linear layer's input feeds
 * the forward GEMM as-is and the weight-gradient GEMM transposed; its output gradient feeds dX and dW likewise). */
int nt_split_bf16_both(const float* x, void* hi, void* lo, void* hi_t, void* lo_t, int rows, int cols, int ld_t);
/* Producer-side splits: ops whose output feeds a BF16x3 GEMM can emit the bf16 hi / lo planes of that output
 * themselves (y_hi / y_lo, dx_hi / dx_lo ...: dense [rows][cols] planes, NULL = not wanted), so no separate split
 * pass has to read the fp32 tensor back from HBM.
 * fclayer.forward (net/fullconnect.py:99-106) from split x [M][K] and split w^T [N][K]; epilogue as nt_linear_fwd. */
int nt_linear_fwd_bf16x3(const void* x_hi, const void* x_lo, const void* wt_hi, const void* wt_lo, const float* b, float* y,
                         int M, int K, int N, int act, float* pre, const float* residual, void* y_hi, void* y_lo);
/* fclayer.backward input half (:114) from split dy [M][N] and split w [K][N]; epilogue as nt_linear_bwd_input. */
int nt_linear_bwd_input_bf16x3(const void* dy_hi, const void* dy_lo, const void* w_hi, const void* w_lo, float* dx, int M, int K,
                               int N, int act, const float* pre, const float* addend, void* dx_hi, void* dx_lo);
/* fclayer.backward parameter half (:118-125) from transposed splits x^T [K][ld_t], dy^T [N][ld_t] (ld_t >= M), or, with
 * ld_t == 0, from the UNtransposed splits x [M][K], dy [M][N] (MN-major tensor-core operands: no transposed copies);
 * db (may be NULL) += colsum(dy) from the fp32 dy.  Accumulates. */
int nt_linear_bwd_params_bf16x3(const void* xt_hi, const void* xt_lo, const void* dyt_hi, const void* dyt_lo, int ld_t,
                                const float* dy, float* dw, float* db, int M, int K, int N);

/* ---------------------------------------------------------------- LayerNorm (net/layernorm.py) */
/* layer_norm.forward, net/layernorm.py:88-114: rows of width D, biased variance, eps inside sqrt.
 * mean/rstd (length M) are kept for backward.  post_add (may be NULL) is a [period,D] table added
 * after the affine (Position_learnable.forward, Position_add.py:20-23, fused). */
int nt_layernorm_fwd(const float* x, const float* gamma, const float* beta, float* y,
                     float* mean, float* rstd, int M, int D, float eps,
                     const float* post_add, int period);
/* layer_norm.backward, net/layernorm.py:116-135: dx = three-term formula (+ addend if non-NULL);
 * dgamma[D] += sum xhat*dy ; dbeta[D] += sum dy.  Accumulates into dgamma/dbeta. */
int nt_layernorm_bwd(const float* dy, const float* x, const float* mean, const float* rstd,
                     const float* gamma, float* dx, float* dgamma, float* dbeta,
                     int M, int D, const float* addend);
/* the same with the output also written as bf16 hi / lo planes (operand of the next BF16x3 GEMM) */
int nt_layernorm_fwd_split(const float* x, const float* gamma, const float* beta, float* y, float* mean, float* rstd,
                           int M, int D, float eps, const float* post_add, int period, void* y_hi, void* y_lo);
int nt_layernorm_bwd_split(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma,
                           float* dx, float* dgamma, float* dbeta, int M, int D, const float* addend, void* dx_hi, void* dx_lo);

/* ---------------------------------------------------------------- activations (net/activation.py) */
int nt_gelu_fwd(const float* x, float* y, size_t n);                       /* :139-142 */
int nt_gelu_bwd(const float* dy, const float* x, float* dx, size_t n);     /* :144-148 */
int nt_relu_fwd(const float* x, float* y, size_t n);                       /* :11-14  */
int nt_relu_bwd(const float* dy, const float* x_pre, float* dx, size_t n); /* :16-17  */
int nt_softmax_fwd(const float* x, float* p, int rows, int cols);          /* :76-84, axis=-1 */
int nt_softmax_bwd(const float* dp, const float* p, float* dx, int rows, int cols); /* :86-129 closed form */
int nt_add(const float* a, const float* b, float* out, size_t n);          /* residual adds */
int nt_sigmoid_fwd(const float* x, float* y, size_t n);                    /* :27-30  */
int nt_sigmoid_bwd(const float* dy, const float* y, float* dx, size_t n);  /* :32-33, y = forward output */
int nt_tanh_fwd(const float* x, float* y, size_t n);                       /* :43-47  */
int nt_tanh_bwd(const float* dy, const float* y, float* dx, size_t n);     /* :49-50, y = forward output */
int nt_silu_fwd(const float* x, float* y, size_t n);                       /* :59-63  */
int nt_silu_bwd(const float* dy, const float* x, float* dx, size_t n);     /* :65-67, x = forward input */
int nt_add_rows(const float* x, const float* table, float* out, int rows, int period, int D); /* Position_add.py:20-23 */
int nt_scale(float* x, float alpha, size_t n);

/* ---------------------------------------------------------------- attention core
 * attention.py:31-46,63-84 and gpt/attdecoderblock.py:52-69,79-99 (the per-sample/per-head loops).
 * qkv [B,N,3*H*dh], columns ordered [q,k,v][head][dh] (attention.py:24).
 * mask_kind: 0 none, 1 causal (-inf above the diagonal, gpt_train_potry3000.py:84-91),
 *            2 dense additive fp32 mask[N,N] (added AFTER the 1/sqrt(dh) scale, attention.py:37-39).
 * P [B,H,N,N] receives the softmax (the reference's atg__), S (may be NULL) the pre-mask scaled
 * scores (att_col, attdecoderblock.py:58-59).  out[B,N,H*dh] = concat_h(P V) (+ residual if non-NULL). */
#define NT_MASK_NONE 0
#define NT_MASK_CAUSAL 1
#define NT_MASK_DENSE 2
int nt_attention_fwd(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S,
                     int B, int N, int H, int dh, const float* residual);
/* dqkv[B,N,3*H*dh] from dout[B,N,H*dh]; scratch must hold B*H*N*N floats (dS) when N > 64.  With dh == 64 in GEMM
 * mode 1 a non-NULL scratch of >= B*H*N floats selects the BF16x3 kernels (attention_bf16.cu: dS is recomputed,
 * only the softmax-backward row sums are parked in scratch).
 * mask_kind is the forward's mask kind: NT_MASK_CAUSAL lets the kernels skip the upper-triangular
 * blocks (P is exactly 0 there); any other value processes the full matrix. */
int nt_attention_bwd(const float* dout, const float* qkv, const float* P, float* dqkv, float* scratch,
                     int B, int N, int H, int dh, int mask_kind);
/* BF16x3 kernels only (dh == 64, GEMM mode 1): out / dqkv additionally as bf16 hi / lo planes for the Linear that
 * consumes them (fc_out forward, qkvfc backward; both may be NULL).  att_out (may be NULL): the forward's attention
 * output [B,N,H*dh] WITHOUT a residual; when given, the softmax-backward row sums sum_j dP_ij P_ij are taken as
 * dout_i . att_out_i (same quantity, attention.py:75) instead of a sweep over the key blocks. */
int nt_attention_fwd_split(const float* qkv, const float* mask, int mask_kind, float* out, float* P, float* S,
                           int B, int N, int H, int dh, const float* residual, void* out_hi, void* out_lo);
int nt_attention_bwd_split(const float* dout, const float* qkv, const float* P, float* dqkv, float* scratch,
                           int B, int N, int H, int dh, int mask_kind, void* dqkv_hi, void* dqkv_lo, const float* att_out);
/* attention.py:31-46 / gpt/attdecoderblock.py:52-69 on the tcgen05 kernel (dh == 64, 64 < N <= 256, no dense mask) with
 * qkv ALSO given as row-major bf16 hi / lo planes [B*N][3*H*dh] (what nt_linear_fwd_bf16x3 writes for the QKV Linear):
 * q, k, v tiles reach shared memory by TMA in tensor-core operand layout.  lse (may be NULL): [B,H,N] row log-sum-exp
 * in the log2 domain, for nt_attention_bwd_planes.  P may be NULL when lse is given (the probabilities are then not
 * materialised: 100 MB per call at the poetry3000 shape that only the layer API's atg__ needs).  *taken = 0 and nothing done if the shape is not covered (call
 * nt_attention_fwd_split instead). */
int nt_attention_fwd_planes(const float* qkv, const void* qkv_hi, const void* qkv_lo, int mask_kind, float* out, float* P,
                            int B, int N, int H, int dh, const float* residual, void* out_hi, void* out_lo, float* lse, int* taken);
/* attention.py:63-84 / gpt/attdecoderblock.py:79-99 on tcgen05 for the same shapes: dout and qkv as bf16 hi / lo planes
 * (TMA operands), att_out = the forward's output WITHOUT a residual (row sums of dP o P = dout . att_out), lse [B,H,N] =
 * the row log-sum-exp the forward wrote (the probabilities are recomputed from q and k, P is not read).  Writes dqkv (+ its
 * planes when given).  *taken = 0 and nothing done if the shape is not covered. */
int nt_attention_bwd_planes(const float* dout, const void* dout_hi, const void* dout_lo, const void* qkv_hi, const void* qkv_lo,
                            const float* att_out, const float* lse, float* dqkv, void* dqkv_hi, void* dqkv_lo,
                            int B, int N, int H, int dh, int mask_kind, int* taken);

/* KV-cache decode.  The reference generates by re-running the whole window for every new token
 * (gpt/gpt_predict_poetrythree.py:108-121, gpt/gpt_train_potry3000.py:300-342); while the window is still growing the
 * keys / values of earlier tokens are unchanged, so they are cached per block.
 * nt_kv_append: copy the k and v columns of qkv [B,T,3D] into caches [B,Tmax,D] at rows pos0..pos0+T-1.
 * nt_attention_decode: out[b, h*dh..] = softmax(q K^T / sqrt(dh)) V for the single new query qkv_new [B,3D] over cache
 * rows 0..len-1 (its own k, v already appended). */
int nt_kv_append(const float* qkv, float* kcache, float* vcache, int B, int T, int Tmax, int pos0, int D);
int nt_attention_decode(const float* qkv_new, const float* kcache, const float* vcache, float* out, int B, int Tmax, int len,
                        int H, int dh);
/* The same steps with the position read from device memory (*d_pos = index of the new token; len = *d_pos + 1), so ONE
 * captured CUDA graph replays for every decode step: nt_decode_embed (token row + position row *d_pos,
 * PatchEmbed.py:132-138), nt_kv_append_at (T = 1), nt_attention_decode_at, nt_argmax_next (greedy choice,
 * gpt_predict_poetrythree.py:115-117 with np.argmax: tok[b] and hist[b, *d_pos + 1] = first maximum of logits[b, :cols]),
 * then nt_increment(d_pos). */
int nt_decode_embed(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int D, int Tmax,
                    const int32_t* d_pos, int num_embeddings);
int nt_kv_append_at(const float* qkv, float* kcache, float* vcache, int B, int Tmax, int D, const int32_t* d_pos);
int nt_attention_decode_at(const float* qkv_new, float* kcache, float* vcache, float* out, int B, int Tmax, int H,
                           int dh, const int32_t* d_pos, int append);   /* append != 0: nt_kv_append_at folded in */
int nt_argmax_next(const float* logits, int ld, int cols, int B, int32_t* tok, int32_t* hist, int hist_ld, const int32_t* d_pos);

/* ---------------------------------------------------------------- fused ViT encoder blocks
 * attention_layer.forward / backward (attention.py:20-55, 57-89) for a whole stack of blocks in ONE launch
 * each way: one CTA per image walks every block (LN -> qkv -> attention -> +res -> LN1 -> fc0 -> GELU ->
 * fc1 -> +res, and the hand-derived adjoint in reverse).  Only the launch-bound shapes are supported
 * (N <= 64 tokens, (D,H) in {(96,4),(96,6),(64,4),(64,2),(32,2)}, no mask); contractions are BF16x3
 * error-compensated mma (fp32-class accuracy).  All members are DEVICE pointers; the array itself is HOST memory
 * and is copied into the kernel arguments (<= 16 blocks per call). */
typedef struct NtVitBlock {
  const float *ln_g, *ln_b;        /* norm:  gamma, beta [D]                      attention.py:12 */
  const float *wqkv, *bqkv;        /* qkvfc: [D,3D], [3D], columns [q,k,v][head][dh]  :14,:24 */
  const float *ln1_g, *ln1_b;      /* norm1                                       :13 */
  const float *w0, *b0;            /* fc0: [D,2D], [2D]                           :15 */
  const float *w1, *b1;            /* fc1: [2D,D], [D]                            :16 */
  float *d_ln_g, *d_ln_b, *d_wqkv, *d_bqkv, *d_ln1_g, *d_ln1_b, *d_w0, *d_b0, *d_w1, *d_b1;  /* accumulate (+=) */
  float *qkv;                      /* [B,N,3D]   kept for backward                :23 */
  float *P;                        /* [B,H,N,N]  softmax, the reference's atg__   :41 */
  float *h;                        /* [B,N,D]    input1 = x + attention           :46 */
  float *dgelu;                    /* [B,N,2D]   GELU'(fc0 output), the factor GELU.backward applies (net/activation.py:144-148) */
  float *out;                      /* [B,N,D]    block output                     :53 */
  void *save;                      /* B * nt_vit_blocks_save_bytes(D): MMA operands of the backward exactly as the
                                      forward held them in shared memory (bf16 hi/lo split LN(x), q|k|v, LN1(h),
                                      GELU(pre)) + LayerNorm statistics; moved with 1-D TMA bulk copies */
  void *wfrag;                     /* scratch, nt_vit_blocks_wfrag_bytes(D): forward fills it with the
                                      bf16 hi/lo mma-fragment-ordered copies of the three weight matrices,
                                      backward of the same step re-uses it */
} NtVitBlock;
int nt_vit_blocks_supported(int N, int D, int H, int* ok);
int nt_vit_blocks_wfrag_bytes(int D, size_t* bytes);
int nt_vit_blocks_save_bytes(int D, size_t* bytes_per_image);
/* Re-split the CURRENT weights of the blocks into wfrag (one small launch).  Needed once after every parameter
 * update; a captured step issues it on the side branch so it overlaps the patch embedding. */
int nt_vit_blocks_prepare(const NtVitBlock* blocks, int nblocks, int D, int H);
/* x [B,N,D] is the input of blocks[0]; blocks[i].out feeds blocks[i+1].  prepared != 0: the caller already ran
 * nt_vit_blocks_prepare for the current weights; 0: the call does it first. */
int nt_vit_blocks_fwd(const NtVitBlock* blocks, int nblocks, const float* x, int B, int N, int D, int H, int prepared);
/* dout [B,N,D] = gradient w.r.t. blocks[nblocks-1].out; dx [B,N,D] receives the gradient w.r.t. x;
 * scratch holds 2*B*N*D floats.  Parameter gradients accumulate into the d_* members. */
int nt_vit_blocks_bwd(const NtVitBlock* blocks, int nblocks, const float* x, const float* dout, float* dx,
                      float* scratch, int B, int N, int D, int H);

/* The same stack on the 5th-generation tensor cores (csrc/vit_block_tc.cu): tcgen05.mma with the accumulators in TMEM,
 * operands as swizzled bf16 hi/lo planes in shared memory, weights streamed by bulk copies.  Shape: N <= 64, D = 96, H = 4
 * (the README configuration).  NtVitBlock is interpreted slightly differently: `dgelu` receives the fc0 PRE-activation
 * [B,N,2D] (GELU and GELU' are recomputed in the backward), `P` may be NULL (the softmax is then not materialised; the
 * backward recomputes it from q, k), `save` is unused, `wfrag` holds nt_vit_tc_wimg_bytes(D) bytes of weight tiles. */
int nt_vit_tc_supported(int N, int D, int H, int* ok);
int nt_vit_tc_wimg_bytes(int D, size_t* bytes);
int nt_attention_tcl_debug(long long* stamps);   /* development aid: [8][2][16] clock64 phase stamps of CTA 0 of the tcgen05 attention forward (64 < N <= 256), NULL = off */
int nt_attention_tcl_bwd_debug(long long* stamps);   /* the same for the backward: [8 pairs][2][16] */
int nt_vit_tc_debug(long long* stamps);   /* development aid: [2][16][16] clock64 phase stamps of CTA 0, NULL = off */
int nt_vit_tc_prepare(const NtVitBlock* blocks, int nblocks, int D, int H);
int nt_vit_tc_fwd(const NtVitBlock* blocks, int nblocks, const float* x, int B, int N, int D, int H, int prepared);
/* backward of the same stack (needs the forward's qkv / h / pre-activation / block outputs and the weight tiles of the
 * same step); dx [B,N,D] doubles as scratch during the call */
int nt_vit_tc_bwd(const NtVitBlock* blocks, int nblocks, const float* x, const float* dout, float* dx, int B, int N, int D, int H);

/* ---------------------------------------------------------------- embeds */
/* PatchEmbed_flatten.forward patch loops, PatchEmbed.py:22-36: (B,C,Hi,Wi) -> (B, n_patch^2, C*h*w) */
int nt_patchify(const float* images, float* out, int B, int C, int Hi, int Wi, int n_patch);

/* Batch construction on the device: the data set stays resident, a step's batch is one gather.
 * nt_gather_images_u8: out[b, :] = lut256[src[perm[first + b], :]] — the shuffled, normalised slice of
 *   transformer_of_image.py:99-100,128-137 (lut256[v] = float32(v / 255.0), built by the host so values are bit-identical);
 * nt_gather_i32: out[b] = src[perm[first + b]] (labels);
 * nt_gather_windows: inputs[b, t] = corpus[start[b] + t], labels[b, t] = corpus[start[b] + t + 1]
 *   (getinputs, gpt/gpt_train_potry3000.py:106-133). */
int nt_gather_images_u8(const unsigned char* src, const int32_t* perm, int first, const float* lut256, float* out, int B,
                        int pixels);
int nt_gather_i32(const int32_t* src, const int32_t* perm, int first, int32_t* out, int B);
int nt_gather_windows(const int32_t* corpus, const int32_t* start, int32_t* inputs, int32_t* labels, int B, int T);

/* convolution_layer (net/Convolution.py:37-248) = im2col + the Linear entry points above.
 * nt_im2col: col[(n, oh, ow)][(c, kh, kw)] of the zero-padded input (:66-84, :114-119); OH = (H + 2 PH - KH) / SH + 1.
 * nt_col2im: its adjoint, dx[N,C,H,W] from dcol (the input gradient of :196-223), written (not accumulated).
 * nt_transpose_last2: y[b][c][r] = x[b][r][c] (NHWC <-> NCHW, kernel tensor <-> Linear weight). */
int nt_im2col(const float* x, float* col, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW);
int nt_col2im(const float* dcol, float* dx, int N, int C, int H, int W, int KH, int KW, int SH, int SW, int PH, int PW);
int nt_transpose_last2(const float* x, float* y, int batch, int rows, int cols);
/* Position_Embedding.forward, PatchEmbed.py:132-138 + Embedding_layer.forward net/Embedding.py:50-55:
 * out[b,t,:] = etok[idx[b,t],:] + epos[t,:]  (epos may be NULL). */
int nt_embedding_fwd(const int32_t* idx, const float* etok, const float* epos, float* out, int B, int T, int D,
                     int num_embeddings);
/* Position_Embedding.backward, PatchEmbed.py:140-144 + Embedding_layer.backward net/Embedding.py:57-70:
 * detok[idx[b,t],:] += dy[b,t,:] ; depos[t,:] += sum_b dy[b,t,:]  (either may be NULL). Accumulates. */
int nt_embedding_bwd(const int32_t* idx, const float* dy, float* detok, float* depos, int B, int T, int D,
                     int num_embeddings);
/* Index errors: ids in [-num_embeddings, 0) wrap like NumPy's negative indexing (Embedding.py:46 `self.params[inputs]`);
 * ids outside [-num_embeddings, num_embeddings) — the reference raises IndexError — never touch memory: the lookup
 * yields zeros, the scatter is skipped, and the NEXT nt_sync / nt_d2h returns non-zero with the reason in
 * nt_last_error().  nt_cross_entropy does the same for an int label outside [0, cols). */

/* ---------------------------------------------------------------- loss (net/loss.py:46-51) */
/* cross_entropy_loss: p = softmax(logits), loss = -sum(label*log(p+1e-10))/n_norm, grad = (p-label)/n_norm.
 * labels are int32 class indices when onehot == NULL, else the dense label matrix [rows,cols].
 * argmax (may be NULL) receives np.argmax(p, -1) with first-index tie-break
 * (transformer_of_image.py:151).  row_loss is a [rows] scratch; loss a device scalar.
 * n_norm is the GLOBAL row count under data parallelism. */
int nt_cross_entropy(const float* logits, const int32_t* labels, const float* onehot, float n_norm,
                     float* loss, float* row_loss, float* grad, float* prob, int32_t* argmax,
                     int rows, int cols);

/* ---------------------------------------------------------------- optimiser */
/* fclayer.update SGD branch + setzero, net/fullconnect.py:150-153,129-132: p -= lr*g ; g = 0 if zero_grad.
 * lr_dev (device float*, may be NULL) overrides lr so a captured graph can follow a schedule. */
/* net/loss.py:53-57 binary_cross_entropy_logit_loss: loss = -mean(y log(s + 1e-10) + (1 - y) log(1 - s + 1e-10)),
 * grad = (s - y) / n, sigmoid = s (grad / sigmoid may be NULL);  net/loss.py:68-71 mean_square_loss: loss = mean((p - y)^2),
 * grad = 2 (p - y) / n.  n = number of elements. */
int nt_bce_logit_loss(const float* logits, const float* label, float* loss, float* grad, float* sigmoid, size_t n);
int nt_mse_loss(const float* predict, const float* label, float* loss, float* grad, size_t n);

int nt_sgd(float* p, float* g, size_t n, float lr, const float* lr_dev, int zero_grad);
/* fclayer.update Adam branch, net/fullconnect.py:137-149 (sqrt on both bias corrections, eps 1e-8
 * outside the sqrt).  t_dev (device int*, may be NULL) overrides t. */
int nt_adam_star(float* p, float* g, float* m, float* v, size_t n, float lr, const float* lr_dev,
                 int t, const int32_t* t_dev, int zero_grad);
int nt_increment(int32_t* counter);      /* t += 1 on device (fullconnect.py:149) */

/* ---------------------------------------------------------------- phase programs (launch-bound GPT shapes) */
/* The fused decoder block of SURVEY.md 8(b) for the shapes where a step is a chain of <= 16-row ops (gpt_character:
 * 10 token rows; one KV-cache decode step): gpt/attdecoderblock.py:31-111 forward and backward, the stem, head and
 * loss around it (PatchEmbed.py:132-146, net/layernorm.py, gpt/classify_gpt.py:20-28, net/loss.py:46-51) as ONE kernel.
 * Between nt_fuse_begin and nt_fuse_end the entry points nt_linear_fwd / _ln_skinny / _bwd_input / _bwd_params,
 * nt_layernorm_fwd / _bwd, nt_attention_fwd / _bwd (N <= 32), nt_attention_decode_at, nt_cross_entropy,
 * nt_embedding_fwd / _bwd, nt_decode_embed, nt_argmax_next and nt_increment do not launch: each call appends one phase
 * to a host-side program.  nt_fuse_end — or any call the program cannot express, which then runs as usual — launches
 * the program as a single thread-block cluster that walks the phases in order, every CTA on a slice of each phase, with
 * a cluster barrier only where a phase touches what an earlier one wrote.  Results are those of the per-op kernels to
 * fp32 rounding (exact fp32 FMAs).  Regions nest; capturable like every other call. */
int nt_fuse_begin(void);
int nt_fuse_end(void);
/* attdecoderblock_layer.forward / backward (gpt/attdecoderblock.py:31-75, 77-111) as ONE call each: the block's op sequence
 * — LN, FC0 + ReLU, FC1 + x, LN1, QKV, masked attention, FCout + h, and its reverse — issued inside a fuse region.  With
 * B*T <= 16 rows and T <= 32 that is one phase-program launch each way; any other shape runs the per-op kernels of the
 * entry points above in the same order.  All members are DEVICE pointers; the struct itself is HOST memory.
 * FFN width is 4 D (attdecoderblock.py:23-24), qkv columns are [q,k,v][head][dh] (:52-55). */
typedef struct NtGptBlock {
  const float *ln_g, *ln_b;        /* norm:   gamma, beta [D]                          attdecoderblock.py:19 */
  const float *w0, *b0;            /* fc0:    [D,4D], [4D]                             :23 */
  const float *w1, *b1;            /* fc1:    [4D,D], [D]                              :24 */
  const float *ln1_g, *ln1_b;      /* norm1                                            :20 */
  const float *wqkv, *bqkv;        /* qkvfc:  [D,3D], [3D]                             :21 */
  const float *wo, *bo;            /* fc_out: [D,D], [D]                               :25 */
  float *d_ln_g, *d_ln_b, *d_w0, *d_b0, *d_w1, *d_b1, *d_ln1_g, *d_ln1_b, *d_wqkv, *d_bqkv, *d_wo, *d_bo;   /* accumulate (+=) */
  /* written by forward, read by backward */
  float *out0, *mean0, *rstd0;     /* LN(x) [M,D] and its statistics [M]               :37 */
  float *pre2, *out2;              /* fc0 pre-activation and ReLU output [M,4D]        :38-41 */
  float *out6;                     /* h = x + fc1(...) [M,D]                           :45 */
  float *outnorm, *mean1, *rstd1;  /* LN1(h)                                           :47 */
  float *qkv;                      /* [B,T,3D]                                         :52 */
  float *P;                        /* [B,H,T,T] softmax, the reference's atg__         :64 */
  float *rkk;                      /* attention output before fc_out [M,D]             :68 */
} NtGptBlock;
/* y [B,T,D] = block(x [B,T,D]); mask (may be NULL) is the dense additive [T,T] matrix for mask_kind NT_MASK_DENSE */
int nt_gpt_block_fwd(const NtGptBlock* blk, const float* x, float* y, int B, int T, int D, int H, const float* mask, int mask_kind);
/* dx [B,T,D] from dy [B,T,D]; parameter gradients accumulate into the d_* members; workspace holds
 * nt_gpt_block_bwd_workspace(B,T,D,H) floats of device memory and must differ from every other argument */
int nt_gpt_block_bwd_workspace(int B, int T, int D, int H, size_t* floats);
int nt_gpt_block_bwd(const NtGptBlock* blk, const float* x, const float* dy, float* dx, float* workspace, int B, int T, int D,
                     int H, int mask_kind);
/* totals since start-up: program launches, phases in them, cluster barriers in them; CTAs per cluster (0 = none yet) */
int nt_fuse_stats(long long* launches, long long* phases, long long* barriers, int* cluster_size);
/* development aid (tools/phase_stamps.py): programs launched after this call write the SM clock at the start of every
 * phase, at the end of its body and behind its barrier into dev_int64_buf [phase][CTA][3] (device memory, >= 64 x 16 x 3
 * int64; NULL switches it off again) */
int nt_fuse_debug_stamps(void* dev_int64_buf);

/* ---------------------------------------------------------------- data parallel */
/* One NCCL communicator per process (one process per GPU).  uid is the 128-byte ncclUniqueId. */
int nt_dp_unique_id(void* uid128);
int nt_dp_init(const void* uid128, int rank, int world);
int nt_dp_allreduce_sum(float* buf, size_t n);   /* in place, on the comm stream, ordered after compute */
int nt_dp_join(void);                            /* compute stream waits for outstanding all-reduces */
/* replicas must start identical: rank `root`'s buffer (n 4-byte words) overwrites every other rank's, on the compute stream */
int nt_dp_broadcast(void* buf, size_t n, int root);
int nt_dp_shutdown(void);
/* Small-message all-reduce over NVLink peer memory (csrc/dp_nccl.cu): every rank maps every peer's gradient arena through
 * CUDA IPC and ONE kernel per rank sums in place (two-shot: rank r reduces slice r of all arenas and stores it back into
 * all of them), bracketed by flag barriers in peer memory.  Used instead of the NCCL call when the range is small enough to
 * be latency-bound (the README ViT: 3.4 MB).  handles are 64-byte cudaIpcMemHandle_t, world of them, rank-ordered. */
int nt_p2p_flags_alloc(void** dptr);
int nt_p2p_ipc_handle(void* dptr, void* handle64);
int nt_p2p_setup(int rank, int world, float* my_grads, const void* grad_handles, void* my_flags, const void* flag_handles);
int nt_p2p_allreduce_sum(size_t off, size_t n, int channel);
int nt_p2p_shutdown(void);

#ifdef __cplusplus
}
#endif
#endif
