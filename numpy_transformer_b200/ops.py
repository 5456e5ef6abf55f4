"""Thin functional wrappers: DeviceArray in, DeviceArray out, one C-ABI call each (see include/ntb200.h).
Nothing here touches the host copy of a tensor, so every function is graph-capturable."""
import ctypes

import numpy as np
from ._lib import lib
from .device import DeviceArray, as_device, ptr_or_null, _prod, get_gemm_mode

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2
MASK_NONE, MASK_CAUSAL, MASK_DENSE = 0, 1, 2
LN_EPS = 1e-5        # net/layernorm.py:86
_last = {}           # which engine the last call of an op took (tests assert on it)


def bf16x3_ok(M, K, N):
    """Large Linear shapes in GEMM mode 1 run on the BF16x3 tcgen05 engine (operands split to bf16 hi/lo in HBM,
    csrc/split_bf16.cu): fp32-class accuracy at twice the tensor rate of the in-kernel TF32 split.  Small shapes are
    latency-bound and stay on the TF32x3 kernel (no extra split launches).  NTB200_BF16X3=0 disables it."""
    import os
    if os.environ.get("NTB200_BF16X3", "1") == "0" or get_gemm_mode() != 1:
        return False
    return M >= int(os.environ.get("NTB200_BF16X3_MIN_ROWS", "2048")) and K % 8 == 0 and N % 8 == 0 and K >= 32 and N >= 32


def want_split(rows, cols):
    """A producer whose [rows, cols] output feeds a Linear should also emit the bf16 hi / lo planes (the consumer will
    take the BF16x3 engine): saves the separate split pass over the fp32 tensor.  NTB200_PRODUCER_SPLIT=0 disables it."""
    import os
    return os.environ.get("NTB200_PRODUCER_SPLIT", "1") != "0" and bf16x3_ok(rows, cols, cols)


def new_planes(n_elems):
    return DeviceArray((n_elems // 2,)), DeviceArray((n_elems // 2,))          # two bf16 per fp32 word


def split_bf16(x, rows, cols, transpose=False):
    """-> (hi, lo, pitch): bf16 planes of x [rows, cols] (or of its transpose, pitch padded to a multiple of 8).
    Planes attached by the producer of x (x._split) are reused."""
    if not transpose and getattr(x, "_split", None) is not None:
        return x._split[0], x._split[1], cols
    ld = cols if not transpose else (rows + 7) // 8 * 8
    out_rows = cols if transpose else rows
    hi = DeviceArray((out_rows * ld // 2,))          # two bf16 per fp32 word
    lo = DeviceArray((out_rows * ld // 2,))
    lib.nt_split_bf16(x.ptr, hi.ptr, lo.ptr, int(rows), int(cols), int(ld), int(bool(transpose)))
    return hi, lo, ld


def split_bf16_both(x, rows, cols):
    """-> ((hi, lo), (hi_t, lo_t, pitch)): both layouts of x [rows, cols] from one pass over the fp32 tensor."""
    ld = (rows + 7) // 8 * 8
    hi, lo = DeviceArray((rows * cols // 2,)), DeviceArray((rows * cols // 2,))
    hit, lot = DeviceArray((cols * ld // 2,)), DeviceArray((cols * ld // 2,))
    lib.nt_split_bf16_both(x.ptr, hi.ptr, lo.ptr, hit.ptr, lot.ptr, int(rows), int(cols), int(ld))
    return (hi, lo), (hit, lot, ld)


def _mn_major():
    """dW reads the UNtransposed splits as MN-major tensor-core operands (default); NTB200_BF16X3_MN=0 restores the
    transposed copies."""
    import os
    return os.environ.get("NTB200_BF16X3_MN", "1") != "0"


def linear_fwd(x, w, b=None, act=ACT_NONE, want_pre=False, residual=None, keep_xt=None, split_out=False, keep_w=None):
    """keep_xt: a list; on the BF16x3 path the split of x that the weight-gradient GEMM will need (the forward's own
    operand when dW reads MN-major operands, else the transposed split made in the same pass) is appended to it.
    keep_w: a list; the UNtransposed split of w (the operand of the dX GEMM) is made in the same pass as the forward's
    transposed one and appended, so the backward does not read w again."""
    K, N = w.shape
    assert x.shape[-1] == K, "linear_fwd: input feature %d != %d" % (x.shape[-1], K)
    M = x.size // K
    y = DeviceArray(x.shape[:-1] + (N,), host_dtype=x.host_dtype)
    pre = y.empty_like() if want_pre else None
    if bf16x3_ok(M, K, N):
        if keep_xt is not None and not _mn_major():
            (xh, xl), xt = split_bf16_both(x, M, K)
            keep_xt.append(xt)
        else:
            xh, xl, _ = split_bf16(x, M, K)
            if keep_xt is not None:
                keep_xt.append((xh, xl, 0))             # pitch 0 = untransposed
        if keep_w is not None:
            wu, (wh, wl, _) = split_bf16_both(w, K, N)             # w [K][N] for dX and w^T [N][K]: one pass over w
            keep_w.append(wu)
        else:
            wh, wl, _ = split_bf16(w, K, N, transpose=True)        # w^T [N][K]
        yh = yl = None
        if split_out and want_split(M, N):
            yh, yl = new_planes(M * N)
            y._split = (yh, yl)
        lib.nt_linear_fwd_bf16x3(xh.ptr, xl.ptr, wh.ptr, wl.ptr, ptr_or_null(b), y.ptr, M, K, N, act, ptr_or_null(pre),
                                 ptr_or_null(residual), ptr_or_null(yh), ptr_or_null(yl))
        return (y, pre) if want_pre else y
    lib.nt_linear_fwd(x.ptr, w.ptr, ptr_or_null(b), y.ptr, M, K, N, act, ptr_or_null(pre), ptr_or_null(residual))
    return (y, pre) if want_pre else y


def linear_bwd_input(dy, w, act=ACT_NONE, pre=None, addend=None, dy_split=None, split_out=False, w_split=None):
    K, N = w.shape
    M = dy.size // N
    dx = DeviceArray(dy.shape[:-1] + (K,), host_dtype=dy.host_dtype)
    if bf16x3_ok(M, N, K):
        dh, dl = dy_split if dy_split is not None else split_bf16(dy, M, N)[:2]
        wh, wl = w_split if w_split is not None else split_bf16(w, K, N)[:2]   # w [K][N]: contraction over N is contiguous
        xh_ = xl_ = None
        if split_out and want_split(M, K):
            xh_, xl_ = new_planes(M * K)
            dx._split = (xh_, xl_)
        lib.nt_linear_bwd_input_bf16x3(dh.ptr, dl.ptr, wh.ptr, wl.ptr, dx.ptr, M, K, N, act, ptr_or_null(pre), ptr_or_null(addend),
                                       ptr_or_null(xh_), ptr_or_null(xl_))
        return dx
    lib.nt_linear_bwd_input(dy.ptr, w.ptr, dx.ptr, M, K, N, act, ptr_or_null(pre), ptr_or_null(addend))
    return dx


def linear_bwd_params(x, dy, dw, db=None, xt_split=None, dyt_split=None):
    K, N = dw.shape
    M = dy.size // N
    assert x.size == M * K, "linear_bwd_params: x %s vs dy %s" % (x.shape, dy.shape)
    if bf16x3_ok(M, K, N):
        if _mn_major():
            # x [M][K] and dy [M][N] as they are: MN-major operands of the contraction over the M rows
            xh, xl = xt_split[:2] if (xt_split is not None and xt_split[2] == 0) else split_bf16(x, M, K)[:2]
            dh, dl = dyt_split[:2] if (dyt_split is not None and dyt_split[2] == 0) else split_bf16(dy, M, N)[:2]
            lib.nt_linear_bwd_params_bf16x3(xh.ptr, xl.ptr, dh.ptr, dl.ptr, 0, dy.ptr, dw.ptr, ptr_or_null(db), M, K, N)
            return
        xh, xl, ld = xt_split if (xt_split is not None and xt_split[2]) else split_bf16(x, M, K, transpose=True)      # x^T [K][M]
        dh, dl, _ = dyt_split if (dyt_split is not None and dyt_split[2]) else split_bf16(dy, M, N, transpose=True)    # dy^T [N][M]
        lib.nt_linear_bwd_params_bf16x3(xh.ptr, xl.ptr, dh.ptr, dl.ptr, ld, dy.ptr, dw.ptr, ptr_or_null(db), M, K, N)
        return
    lib.nt_linear_bwd_params(x.ptr, dy.ptr, dw.ptr, ptr_or_null(db), M, K, N)


def layernorm_fwd(x, gamma, beta, post_add=None, split_out=False):
    D = gamma.size
    M = x.size // D
    y = x.empty_like()
    mean = DeviceArray((M,))
    rstd = DeviceArray((M,))
    period = 0 if post_add is None else post_add.size // D
    if split_out and want_split(M, D):
        yh, yl = new_planes(M * D)
        y._split = (yh, yl)
        lib.nt_layernorm_fwd_split(x.ptr, gamma.ptr, beta.ptr, y.ptr, mean.ptr, rstd.ptr, M, D, LN_EPS,
                                   ptr_or_null(post_add), period, yh.ptr, yl.ptr)
        return y, mean, rstd
    lib.nt_layernorm_fwd(x.ptr, gamma.ptr, beta.ptr, y.ptr, mean.ptr, rstd.ptr, M, D, LN_EPS,
                         ptr_or_null(post_add), period)
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, addend=None, want_dx=True, split_out=False):
    D = gamma.size
    M = x.size // D
    dx = x.empty_like() if want_dx else None
    if (split_out and want_dx and want_split(M, D) and D <= 768 and (dgamma is None) == (dbeta is None)):
        dh, dl = new_planes(M * D)
        dx._split = (dh, dl)
        lib.nt_layernorm_bwd_split(dy.ptr, x.ptr, mean.ptr, rstd.ptr, gamma.ptr, dx.ptr, ptr_or_null(dgamma), ptr_or_null(dbeta),
                                   M, D, ptr_or_null(addend), dh.ptr, dl.ptr)
        return dx
    lib.nt_layernorm_bwd(dy.ptr, x.ptr, mean.ptr, rstd.ptr, gamma.ptr, ptr_or_null(dx), ptr_or_null(dgamma),
                         ptr_or_null(dbeta), M, D, ptr_or_null(addend))
    return dx


def _att_bf16(N, dh):
    import os
    return (dh == 64 and N >= int(os.environ.get("NTB200_ATT_BF16_MIN_N", "16")) and get_gemm_mode() in (1, 2)
            and os.environ.get("NTB200_ATT_BF16", "1") != "0")


def attention_wants_planes(N, dh, mask_kind, want_scores=False):
    """True when attention_fwd would take qkv as bf16 hi / lo planes (the tcgen05 forward for 64 < N <= 256, dh = 64,
    csrc/attention_tcl.cu): the QKV Linear should then be run with split_out=True.  NTB200_ATT_TCL=0 / NTB200_ATT_PLANES=0
    switch the kernel / the TMA operand path off."""
    import os
    return (dh == 64 and 64 < N <= 256 and mask_kind != MASK_DENSE and not want_scores and get_gemm_mode() in (1, 2)
            and os.environ.get("NTB200_ATT_TCL", "1") != "0" and os.environ.get("NTB200_ATT_PLANES", "1") != "0")


_lean = [False]


class lean_attention:
    """Context: attention_fwd may skip materialising P where the backward does not need it (TrainStep; the layer API keeps
    P because the reference exposes it as atg__ / att)."""

    def __enter__(self):
        self.prev, _lean[0] = _lean[0], True

    def __exit__(self, *exc):
        _lean[0] = self.prev


class LazyP:
    """Stand-in for the probabilities [B,H,N,N] when the forward did not write them: carries what the tcgen05 backward takes
    instead (row log-sum-exp, planes of qkv) and recomputes P on the device the first time anybody asks (np.asarray, .ptr)."""

    def __init__(self, qkv, H, mask_kind, lse, planes, host_dtype):
        B, N, _ = qkv.shape
        self.shape, self.host_dtype = (B, H, N, N), host_dtype
        self._qkv, self._H, self._mask_kind = qkv, H, mask_kind
        self._lse, self._qkv_planes, self._dense = lse, planes, None

    def _materialise(self):
        if self._dense is None:
            prev, _lean[0] = _lean[0], False
            try:
                self._dense = attention_fwd(self._qkv, self._H, None, self._mask_kind)[1]
            finally:
                _lean[0] = prev
        return self._dense

    @property
    def ptr(self):
        return self._materialise().ptr

    def numpy(self):
        return self._materialise().numpy()

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self._materialise())
        return a if dtype is None else a.astype(dtype, copy=False)


def attention_fwd(qkv, H, mask=None, mask_kind=MASK_NONE, residual=None, want_scores=False, split_out=False):
    B, N, three_d = qkv.shape
    D = three_d // 3
    out = DeviceArray((B, N, D), host_dtype=qkv.host_dtype)
    import os
    # (residual is None: the GPT block, whose backward hands attention_bwd the residual-free output the tcgen05 backward needs;
    # a caller that fuses a residual here would fall back to the kernels that read P)
    lean = (_lean[0] and not want_scores and residual is None and os.environ.get("NTB200_ATT_TCL_BWD", "1") != "0" and B * N >= 2048
            and attention_wants_planes(N, D // H, mask_kind, want_scores))
    P = None if lean else DeviceArray((B, H, N, N), host_dtype=qkv.host_dtype)
    S = P.empty_like() if want_scores else None
    sp = getattr(qkv, "_split", None)
    if sp is None and B * N >= 2048 and attention_wants_planes(N, D // H, mask_kind, want_scores):
        # the producer did not emit the planes (mode 2: the QKV Linear is a single-pass TF32 GEMM): one split pass over qkv is
        # cheaper than what the TMA-fed kernels save (G1: ~30 us against ~150 us per layer, forward + backward)
        sp = split_bf16(qkv, B * N, 3 * D)[:2]
    if sp is not None and attention_wants_planes(N, D // H, mask_kind, want_scores):
        # qkv came with its bf16 planes (QKV Linear with split_out=True): the tcgen05 kernel takes its operands by TMA
        oh = ol = None
        if split_out and want_split(B * N, D):
            oh, ol = new_planes(B * N * D)
        taken = ctypes.c_int(0)
        lse = DeviceArray((B * H * N,))
        lib.nt_attention_fwd_planes(qkv.ptr, sp[0].ptr, sp[1].ptr, mask_kind, out.ptr, ptr_or_null(P), B, N, H, D // H,
                                    ptr_or_null(residual), ptr_or_null(oh), ptr_or_null(ol), lse.ptr, ctypes.byref(taken))
        if taken.value:
            qkv._split = None
            # what the tcgen05 backward takes instead of P: the planes of qkv and the row log-sum-exp (attention_bwd)
            if P is None:
                P = LazyP(qkv, H, mask_kind, lse, sp, qkv.host_dtype)
            else:
                P._lse, P._qkv_planes = lse, sp
            if oh is not None:
                out._split = (oh, ol)
            return out, P, S
    if P is None:
        P = DeviceArray((B, H, N, N), host_dtype=qkv.host_dtype)
    if split_out and _att_bf16(N, D // H) and want_split(B * N, D):
        oh, ol = new_planes(B * N * D)
        out._split = (oh, ol)
        lib.nt_attention_fwd_split(qkv.ptr, ptr_or_null(mask), mask_kind, out.ptr, P.ptr, ptr_or_null(S), B, N, H, D // H,
                                   ptr_or_null(residual), oh.ptr, ol.ptr)
        return out, P, S
    lib.nt_attention_fwd(qkv.ptr, ptr_or_null(mask), mask_kind, out.ptr, P.ptr, ptr_or_null(S), B, N, H, D // H,
                         ptr_or_null(residual))
    return out, P, S


def attention_bwd(dout, qkv, P, H, mask_kind=MASK_NONE, split_out=False, att_out=None):
    """att_out: the forward's attention output without a residual (lets the long-sequence BF16x3 kernel take the
    softmax-backward row sums as dout . att_out instead of a sweep over the key blocks)."""
    B, N, three_d = qkv.shape
    D = three_d // 3
    dh = D // H
    dqkv = qkv.empty_like()
    dsp = getattr(dout, "_split", None)
    if (dsp is None and getattr(P, "_lse", None) is not None and att_out is not None and B * N >= 2048
            and attention_wants_planes(N, dh, mask_kind)):
        dsp = split_bf16(dout, B * N, D)[:2]          # see attention_fwd
    if (getattr(P, "_lse", None) is not None and dsp is not None and att_out is not None
            and attention_wants_planes(N, dh, mask_kind)):
        # tcgen05 backward (csrc/attention_tcl.cu): dout and qkv as bf16 planes by TMA, P recomputed from the row log-sum-exp
        gh = gl = None
        if split_out and want_split(B * N, 3 * D):
            gh, gl = new_planes(B * N * 3 * D)
        taken = ctypes.c_int(0)
        lib.nt_attention_bwd_planes(dout.ptr, dsp[0].ptr, dsp[1].ptr, P._qkv_planes[0].ptr, P._qkv_planes[1].ptr, att_out.ptr,
                                    P._lse.ptr, dqkv.ptr, ptr_or_null(gh), ptr_or_null(gl), B, N, H, dh, int(mask_kind),
                                    ctypes.byref(taken))
        if taken.value:
            _last["attention_bwd"] = "tcgen05-planes"
            if gh is not None:
                dqkv._split = (gh, gl)
            return dqkv
    _last["attention_bwd"] = "mma.sync"
    if _att_bf16(N, dh):
        scratch = DeviceArray((B * H * N,))             # BF16x3 kernels: row sums of dP o P only
        gh = gl = None
        if split_out and want_split(B * N, 3 * D):
            gh, gl = new_planes(B * N * 3 * D)
            dqkv._split = (gh, gl)
        if gh is not None or att_out is not None:
            lib.nt_attention_bwd_split(dout.ptr, qkv.ptr, P.ptr, dqkv.ptr, scratch.ptr, B, N, H, dh, int(mask_kind),
                                       ptr_or_null(gh), ptr_or_null(gl), ptr_or_null(att_out))
            return dqkv
    else:
        scratch = None if (N <= 64 and dh <= 64 and dh % 4 == 0) else P.empty_like()
    lib.nt_attention_bwd(dout.ptr, qkv.ptr, P.ptr, dqkv.ptr, ptr_or_null(scratch), B, N, H, dh, int(mask_kind))
    return dqkv


def cross_entropy(logits, labels=None, onehot=None, n_norm=None, want_prob=True, want_grad=True, loss_out=None,
                  argmax_out=None):
    rows, cols = logits.shape
    n_norm = float(rows if n_norm is None else n_norm)
    loss = loss_out if loss_out is not None else DeviceArray((), host_dtype=np.float64)
    row_loss = DeviceArray((rows,))
    grad = logits.empty_like() if want_grad else None
    prob = logits.empty_like() if want_prob else None
    amax = argmax_out if argmax_out is not None else DeviceArray((rows,), dtype=np.int32)
    lib.nt_cross_entropy(logits.ptr, ptr_or_null(labels), ptr_or_null(onehot), n_norm, loss.ptr, row_loss.ptr,
                         ptr_or_null(grad), ptr_or_null(prob), amax.ptr, rows, cols)
    if prob is not None:
        prob._argmax = amax
    return loss, grad, prob, amax


def sgd(p, g, lr, zero_grad=False):
    if isinstance(lr, DeviceArray):
        lib.nt_sgd(p.ptr, g.ptr, p.size, 0.0, lr.ptr, int(zero_grad))
    else:
        lib.nt_sgd(p.ptr, g.ptr, p.size, float(lr), None, int(zero_grad))


def adam_star(p, g, m, v, lr, t, zero_grad=False):
    lr_dev = lr.ptr if isinstance(lr, DeviceArray) else None
    t_dev = t.ptr if isinstance(t, DeviceArray) else None
    lib.nt_adam_star(p.ptr, g.ptr, m.ptr, v.ptr, p.size, 0.0 if lr_dev else float(lr), lr_dev,
                     0 if t_dev else int(t), t_dev, int(zero_grad))


def detect_mask(masks, N):
    """Reference callers pass `masks=[]` or a (T,T) additive ndarray (attention.py:38-39).  Returns
    (mask_kind, device_mask_or_None).  A 0 / -inf lower-triangular pattern (gpt_train_potry3000.py:84-91)
    is recognised and handled in-kernel without uploading the matrix."""
    if masks is None:
        return MASK_NONE, None
    if isinstance(masks, str):
        assert masks == "causal"
        return MASK_CAUSAL, None
    if isinstance(masks, DeviceArray):
        return MASK_DENSE, masks
    if len(masks) == 0:
        return MASK_NONE, None
    m = np.asarray(masks)
    assert m.shape == (N, N), "mask must be (%d,%d), got %s" % (N, N, m.shape)
    lower = np.tril(np.ones((N, N), dtype=bool))
    if np.all(m[lower] == 0) and np.all(np.isneginf(m[~lower])):
        return MASK_CAUSAL, None
    if not np.any(m):
        return MASK_NONE, None
    return MASK_DENSE, DeviceArray.from_host(m)
