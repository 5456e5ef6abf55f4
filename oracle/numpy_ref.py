"""NumPy restatement of the reference's per-op math (TEST INFRASTRUCTURE — see oracle/__init__).

Every function is the vectorised closed form of one reference method and cites the
reference file:line it follows (paths relative to the reference root).  All arithmetic runs
in the dtype of the inputs (fp64 in the parity suite).  `*_refstruct` variants keep the
reference's loop structure (per-sample/per-head Python loops, per-row softmax Jacobian) and
are only used to time a "reference-structured" CPU baseline and to cross-check the closed forms.
"""
import numpy as np
from scipy.special import erf

LN_EPS = 1e-5          # net/layernorm.py:86
ADAM_B1, ADAM_B2 = 0.9, 0.999   # net/fullconnect.py:70-72
ADAM_EPS = 10 ** (2 - 10)       # net/fullconnect.py:73


# ----------------------------------------------------------------------------- Linear
def linear_fwd(x, W, b=None):
    """net/fullconnect.py:99-106  y = x @ W (+ b); W is (in, out)."""
    y = np.matmul(x, W)
    if b is not None:
        y = y + b[np.newaxis, :]
    return y


def linear_bwd(dy, x, W):
    """net/fullconnect.py:108-127  dX = dY Wᵀ ; dW = Xᵀ dY ; db = Σ_rows dY (leading dims flattened)."""
    dx = np.matmul(dy, W.T)
    dy2 = dy.reshape(-1, dy.shape[-1])
    x2 = x.reshape(-1, x.shape[-1])
    dW = np.matmul(dy2.T, x2).T
    db = dy2.sum(axis=0)
    return dx, dW, db


# ----------------------------------------------------------------------------- LayerNorm
def layernorm_fwd(x, gamma, beta, eps=LN_EPS):
    """net/layernorm.py:88-114  biased variance over the last axis; returns y, xhat, var."""
    mean = x.mean(axis=-1, keepdims=True)
    var = x.var(axis=-1, keepdims=True)
    xhat = (x - mean) / np.sqrt(var + eps)
    y = xhat * gamma.reshape(-1) + beta.reshape(-1)
    return y, xhat, var


def layernorm_bwd(dy, xhat, var, gamma, eps=LN_EPS):
    """net/layernorm.py:116-135  three-term dx; dγ = Σ x̂·dy, dβ = Σ dy over leading axes."""
    g = gamma.reshape(-1)
    D = xhat.shape[-1]
    rstd = 1.0 / np.sqrt(var + eps)
    p1 = g * dy * rstd
    p2 = rstd / D * np.sum(dy * g, axis=-1, keepdims=True)
    p3 = rstd / D * xhat * np.sum(dy * xhat * g, axis=-1, keepdims=True)
    dx = p1 - p2 - p3
    lead = tuple(range(dy.ndim - 1))
    dgamma = np.sum(xhat * dy, axis=lead)
    dbeta = np.sum(dy, axis=lead)
    return dx, dgamma, dbeta


# ----------------------------------------------------------------------------- activations
def relu_fwd(x):
    """net/activation.py:11-14 (the reference clamps in place; here a new array)."""
    return np.where(x < 0, 0, x)


def relu_bwd(dy, x_pre):
    """net/activation.py:16-17  mask from the saved pre-activation."""
    return (x_pre > 0) * dy


def gelu_fwd(x):
    """net/activation.py:139-142  exact-erf GELU."""
    return 0.5 * x * (1 + erf(x / np.sqrt(2)))


def gelu_bwd(dy, x):
    """net/activation.py:144-148."""
    s = x / np.sqrt(2)
    dx = 0.5 + 0.5 * erf(s) + (0.5 * x * (2 / np.sqrt(np.pi)) * np.exp(-(s ** 2))) / np.sqrt(2)
    return dx * dy


def softmax_fwd(x, axis=-1):
    """net/activation.py:76-84  max-shifted softmax."""
    z = x - np.max(x, axis=axis, keepdims=True)
    e = np.exp(z)
    return e / np.sum(e, axis=axis, keepdims=True)


def softmax_bwd(dp, p):
    """net/activation.py:86-129  d = g·(diag(p) − p pᵀ) per row  ≡  p∘(g − Σ g∘p)."""
    return p * (dp - np.sum(dp * p, axis=-1, keepdims=True))


def softmax_bwd_refstruct(dp, p):
    """net/activation.py:86-91,93-129 with the explicit per-row Jacobian (no ThreadPool)."""
    out = np.zeros(p.shape, dtype=np.float64)
    for i in range(p.shape[0]):
        kk = p[i].reshape(-1, 1)
        jac = np.diagflat(kk) - np.dot(kk, kk.T)
        out[i] = np.dot(dp[i:i + 1], jac)[0]
    return out


# ----------------------------------------------------------------------------- attention core
def attention_fwd(qkv, num_h, mask=None):
    """attention.py:24-46 / gpt/attdecoderblock.py:44-69.

    qkv: (B, N, 3*H*dh) with columns ordered [q,k,v][head][dh] (reshape (B,N,3,H,dh), attention.py:24).
    Returns O (B,N,H*dh) with heads concatenated on features, P (B,H,N,N), S (B,H,N,N) pre-mask scores.
    Scale 1/sqrt(dh) is applied BEFORE the additive mask (attention.py:37-39).
    """
    B, N, three_d = qkv.shape
    D = three_d // 3
    dh = D // num_h
    q5 = qkv.reshape(B, N, 3, num_h, dh)
    q = q5[:, :, 0].transpose(0, 2, 1, 3)
    k = q5[:, :, 1].transpose(0, 2, 1, 3)
    v = q5[:, :, 2].transpose(0, 2, 1, 3)
    S = np.matmul(q, k.transpose(0, 1, 3, 2)) / np.sqrt(dh)
    A = S if mask is None else S + mask
    P = softmax_fwd(A, axis=-1)
    O = np.matmul(P, v).transpose(0, 2, 1, 3).reshape(B, N, D)
    return O, P, S


def attention_bwd(dO, qkv, P, num_h):
    """attention.py:63-84  dP = dO Vᵀ ; dS = softmax'(dP) ; dV = Pᵀ dO ; dQ = dS K/√dh ; dK = dSᵀ Q/√dh."""
    B, N, three_d = qkv.shape
    D = three_d // 3
    dh = D // num_h
    q5 = qkv.reshape(B, N, 3, num_h, dh)
    q = q5[:, :, 0].transpose(0, 2, 1, 3)
    k = q5[:, :, 1].transpose(0, 2, 1, 3)
    v = q5[:, :, 2].transpose(0, 2, 1, 3)
    dOh = dO.reshape(B, N, num_h, dh).transpose(0, 2, 1, 3)
    dP = np.matmul(dOh, v.transpose(0, 1, 3, 2))
    dS = softmax_bwd(dP, P)
    dV = np.matmul(P.transpose(0, 1, 3, 2), dOh)
    dQ = np.matmul(dS, k) / np.sqrt(dh)
    dK = np.matmul(dS.transpose(0, 1, 3, 2), q) / np.sqrt(dh)
    dqkv = np.zeros((B, N, 3, num_h, dh), dtype=dO.dtype)
    dqkv[:, :, 0] = dQ.transpose(0, 2, 1, 3)
    dqkv[:, :, 1] = dK.transpose(0, 2, 1, 3)
    dqkv[:, :, 2] = dV.transpose(0, 2, 1, 3)
    return dqkv.reshape(B, N, 3 * D)


def attention_fwd_refstruct(qkv, num_h, mask=None):
    """attention.py:31-46 with the reference's per-sample / per-head Python loops."""
    B, N, three_d = qkv.shape
    D = three_d // 3
    dh = D // num_h
    q5 = qkv.reshape(B, N, 3, num_h, dh)
    res, Ps = [], [[None] * num_h for _ in range(B)]
    for n in range(B):
        tmp = []
        for i in range(num_h):
            att = np.matmul(q5[n, :, 0, i], q5[n, :, 1, i].T) / np.sqrt(dh)
            if mask is not None:
                att = att + mask
            p = softmax_fwd(att, axis=-1)
            Ps[n][i] = p
            tmp.append(np.matmul(p, q5[n, :, 2, i]))
        res.append(np.concatenate(tmp, axis=-1))
    return np.stack(res), Ps


def attention_bwd_refstruct(dO, qkv, Ps, num_h):
    """attention.py:63-84 with the reference's loops and the explicit softmax Jacobian."""
    B, N, three_d = qkv.shape
    D = three_d // 3
    dh = D // num_h
    q5 = qkv.reshape(B, N, 3, num_h, dh)
    dq5 = np.zeros_like(q5)
    for n in range(B):
        for i in range(num_h):
            q, k, v, p = q5[n, :, 0, i], q5[n, :, 1, i], q5[n, :, 2, i], Ps[n][i]
            d = dO[n][:, dh * i: dh * (i + 1)]
            dp = np.matmul(d, v.T)
            ds = softmax_bwd_refstruct(dp, p)
            dq5[n, :, 2, i] = np.matmul(d.T, p).T
            dq5[n, :, 0, i] = np.matmul(ds, k) / np.sqrt(dh)
            dq5[n, :, 1, i] = np.matmul(ds.T, q) / np.sqrt(dh)
    return dq5.reshape(B, N, 3 * D)


# ----------------------------------------------------------------------------- embeds
def patchify(images, n_patch):
    """PatchEmbed.py:22-36  (B,C,H,W) -> (B, n_patch², C·h·w); patches row-major, (c,h,w) inside."""
    B, C, H, W = images.shape
    hl, wl = H // n_patch, W // n_patch
    x = images[:, :, :hl * n_patch, :wl * n_patch].reshape(B, C, n_patch, hl, n_patch, wl)
    return x.transpose(0, 2, 4, 1, 3, 5).reshape(B, n_patch * n_patch, C * hl * wl)


def posk_table(n_tokens, D):
    """Position_add.py:9-15  sin(i/10000^(j/D)) for even j, cos(i/10000^((j-1)/D)) for odd j."""
    i = np.arange(n_tokens, dtype=np.float64)[:, None]
    j = np.arange(D)[None, :]
    je = np.where(j % 2 == 0, j, j - 1).astype(np.float64)
    ang = i / (10000.0 ** (je / D))
    return np.where(j % 2 == 0, np.sin(ang), np.cos(ang))[None]


def embedding_fwd(idx, E):
    """net/Embedding.py:50-55."""
    return E[idx.reshape(-1)].reshape(list(idx.shape) + [E.shape[1]])


def embedding_bwd(idx, dy, num_embeddings):
    """net/Embedding.py:57-70  scatter-add, duplicates accumulate."""
    D = dy.shape[-1]
    dE = np.zeros((num_embeddings, D), dtype=dy.dtype)
    np.add.at(dE, idx.reshape(-1), dy.reshape(-1, D))
    return dE


# ----------------------------------------------------------------------------- loss / optimiser
def cross_entropy(logits, onehot):
    """net/loss.py:46-51  returns (loss, partial, softmax); N = logits.shape[0]."""
    z = logits - np.max(logits, axis=-1)[..., np.newaxis]
    p = np.exp(z) / np.sum(np.exp(z), axis=-1)[:, np.newaxis]
    loss = -np.sum(onehot * np.log(p + 1e-10)) / logits.shape[0]
    partial = (p - onehot) / logits.shape[0]
    return loss, partial, p


def sgd_update(p, g, lr):
    """net/fullconnect.py:150-153."""
    return p - g * lr


def adam_star_update(p, g, m, v, t, lr):
    """net/fullconnect.py:137-149 — the reference's non-standard Adam: sqrt on BOTH bias corrections,
    eps outside the sqrt.  Returns (p, m, v); the caller advances t."""
    m = ADAM_B1 * m + (1 - ADAM_B1) * g
    v = ADAM_B2 * v + (1 - ADAM_B2) * g ** 2
    mh = m / np.sqrt(1 - ADAM_B1 ** t)
    vh = v / np.sqrt(1 - ADAM_B2 ** t)
    p = p - (mh * lr / (np.sqrt(vh) + ADAM_EPS))
    return p, m, v


# ----------------------------------------------------------------------------- Convolution
def _pair(v):
    return [v, v] if isinstance(v, int) else list(v)


def conv_im2col(x, kernel, stride, padding):
    """net/Convolution.py:66-84 (im2col) on the zero-padded input (:114-119): rows = (n, oh, ow), columns = (c, kh, kw)."""
    (kh, kw), (sh, sw), (ph, pw) = _pair(kernel), _pair(stride), _pair(padding)
    n, c, h, w = x.shape
    oh, ow = (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1
    xp = np.pad(x, ((0, 0), (0, 0), (ph, ph), (pw, pw)))
    col = np.empty((n, oh, ow, c, kh, kw), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            col[:, :, :, :, i, j] = xp[:, :, i:i + sh * oh:sh, j:j + sw * ow:sw].transpose(0, 2, 3, 1)
    return col.reshape(n * oh * ow, c * kh * kw), (n, oh, ow)


def conv_fwd(x, W, b, stride=1, padding=0):
    """net/Convolution.py:104-131  out[n, oc, oh, ow] = sum_{c,kh,kw} xpad[n, c, oh*sh+kh, ow*sw+kw] W[oc, c, kh, kw] (+ b[oc])."""
    oc = W.shape[0]
    col, (n, oh, ow) = conv_im2col(x, W.shape[2:], stride, padding)
    y = col @ W.reshape(oc, -1).T
    if b is not None:
        y = y + b[np.newaxis, :]
    return y.reshape(n, oh, ow, oc).transpose(0, 3, 1, 2)


def conv_bwd(delta, x, W, stride=1, padding=0):
    """net/Convolution.py:163-225 (and the loop form :133-161): db = sum over (n, oh, ow); dW[oc] = sum delta * window;
    dx = scatter of W * delta back through the windows, cropped to the un-padded input."""
    (kh, kw), (sh, sw), (ph, pw) = _pair(W.shape[2:]), _pair(stride), _pair(padding)
    n, c, h, w = x.shape
    oc = W.shape[0]
    col, (_, oh, ow) = conv_im2col(x, (kh, kw), stride, padding)
    d2 = delta.transpose(0, 2, 3, 1).reshape(-1, oc)
    db = d2.sum(axis=0)
    dW = (d2.T @ col).reshape(W.shape)
    dcol = (d2 @ W.reshape(oc, -1)).reshape(n, oh, ow, c, kh, kw)
    dxp = np.zeros((n, c, h + 2 * ph, w + 2 * pw), dtype=x.dtype)
    for i in range(kh):
        for j in range(kw):
            dxp[:, :, i:i + sh * oh:sh, j:j + sw * ow:sw] += dcol[:, :, :, :, i, j].transpose(0, 3, 1, 2)
    return dxp[:, :, ph:ph + h, pw:pw + w], dW, db


# ----------------------------------------------------------------------------- remaining activations / losses
def sigmoid_fwd(x):
    """net/activation.py:27-30."""
    return 1 / (1 + np.exp(-x))


def sigmoid_bwd(dy, y):
    """net/activation.py:32-33 (y = forward output)."""
    return y * (1 - y) * dy


def tanh_fwd(x):
    """net/activation.py:43-47."""
    return np.tanh(x)


def tanh_bwd(dy, y):
    """net/activation.py:49-50 (y = forward output)."""
    return (1 - y ** 2) * dy


def silu_fwd(x):
    """net/activation.py:59-63."""
    return x / (1 + np.exp(-x))


def silu_bwd(dy, x):
    """net/activation.py:65-67: (div + out * e^-x * div) * dy."""
    e = np.exp(-x)
    div = 1 / (1 + e)
    return (div + x * div * e * div) * dy


def bce_logit_loss(predict, label):
    """net/loss.py:53-57."""
    s = 1 / (1 + np.exp(-predict))
    loss = -np.sum(label * np.log(s + 1e-10) + (1 - label) * np.log((1 - s) + 1e-10)) / predict.size
    return loss, (s - label) / predict.size, s


def mse_loss(predict, label):
    """net/loss.py:68-71."""
    return np.sum(np.square(predict - label)) / predict.size, 2 * (predict - label) / predict.size
