"""Model-level oracle: one full training step of the reference's ViT and GPT stacks.

TEST INFRASTRUCTURE (see oracle/__init__).  Parameters, gradients and optimiser state use the
reference's checkpoint tree (attention.py:105-107,
gpt/attdecoderblock.py:129-131, PatchEmbed.py:61-71,154-155, classify.py:50-51):

  ViT : [ [[g,b],[Wpe,bpe]], [], L x [[g,b],[Wqkv,b],[g,b],[W0,b],[W1,b]], [g,b], [[Wc0,b],[Wc1,b]] ]
  GPT : [ [[Etok],[Epos]],       L x [[g,b],[Wqkv,b],[g,b],[W0,b],[W1,b],[Wo,b]], [g,b], [[Wc0,b],[Wc1,b]] ]

so a `save_model()` walk of the reference feeds these functions directly.
"""
from dataclasses import dataclass
import numpy as np
from . import numpy_ref as R


# ----------------------------------------------------------------------------- configs
@dataclass
class ViTConfig:
    """transformer_of_image.py:30-41 (README setting = defaults)."""
    batch: int = 100
    channels: int = 1
    height: int = 28
    width: int = 28
    n_patch: int = 7
    embed_dim: int = 96
    heads: int = 4
    blocks: int = 6
    classes: int = 10
    mlp_ratio: int = 2

    @property
    def tokens(self):
        return self.n_patch ** 2

    @property
    def patch_dim(self):
        return (self.height // self.n_patch) * (self.width // self.n_patch) * self.channels


@dataclass
class GPTConfig:
    """gpt/gpt_train_potry3000.py:153-207 (poetry3000 setting = defaults)."""
    batch: int = 64
    context: int = 256
    embed_dim: int = 384
    heads: int = 6
    blocks: int = 5
    vocab: int = 2350
    mlp_ratio: int = 4


VIT_README = ViTConfig()
VIT_SCALED = ViTConfig(batch=1024, embed_dim=384, heads=6, blocks=12)
GPT_POETRY = GPTConfig()
GPT_SCALED = GPTConfig(batch=512, embed_dim=768, heads=12, blocks=12)
GPT_CHAR = GPTConfig(batch=2, context=5, embed_dim=192, heads=3, blocks=1, vocab=26)


# ----------------------------------------------------------------------------- tree helpers
def tree_map(f, *trees):
    t0 = trees[0]
    if isinstance(t0, (list, tuple)):
        return [tree_map(f, *[t[i] for t in trees]) for i in range(len(t0))]
    return f(*trees)


def tree_leaves(tree):
    if isinstance(tree, (list, tuple)):
        out = []
        for t in tree:
            out.extend(tree_leaves(t))
        return out
    return [tree]


def tree_zeros_like(tree):
    return tree_map(lambda a: np.zeros_like(a), tree)


# ----------------------------------------------------------------------------- initialisers
def _fc_init(rng_uniform, fin, fout):
    """net/fullconnect.py:44-45,53-54  W, b ~ U(±sqrt(1/in)), drawn W first then b."""
    r = np.sqrt(1 / fin)
    W = rng_uniform(-r, r, (fin, fout))
    b = rng_uniform(-r, r, (fout))
    return [W, b]


def _ln_init(D):
    """net/layernorm.py:44-45."""
    return [np.ones(D), np.zeros(D)]


def vit_init(cfg: ViTConfig, seed=0):
    """Draws from the global-style RandomState in the constructor order of
    transformer_of_image.py:56-77 so that `np.random.seed(seed)` + the reference constructors
    give bit-identical parameters (checked by oracle/gen_golden.py)."""
    rs = np.random.RandomState(seed)
    D = cfg.embed_dim
    fc = _fc_init(rs.uniform, cfg.patch_dim, D)
    patch = [_ln_init(D), fc]
    rs.normal(0, 1, (1, cfg.tokens, D))        # Position_add.py:6 draws even when fixed=1
    tree = [patch, []]
    for _ in range(cfg.blocks):                # attention.py:12-16
        qkv = _fc_init(rs.uniform, D, 3 * D)
        fc0 = _fc_init(rs.uniform, D, cfg.mlp_ratio * D)
        fc1 = _fc_init(rs.uniform, cfg.mlp_ratio * D, D)
        tree.append([_ln_init(D), qkv, _ln_init(D), fc0, fc1])
    tree.append(_ln_init(D))
    c0 = _fc_init(rs.uniform, D * cfg.tokens, D)   # classify.py:15
    c1 = _fc_init(rs.uniform, D, cfg.classes)      # classify.py:17
    tree.append([c0, c1])
    return tree


def gpt_init(cfg: GPTConfig, seed=0):
    """Constructor order of gpt/gpt_train_potry3000.py:176-207 (PatchEmbed.py:127-129,
    gpt/attdecoderblock.py:19-25, classify_gpt.py)."""
    rs = np.random.RandomState(seed)
    D = cfg.embed_dim
    etok = rs.normal(0, 1, (cfg.vocab, D))
    epos = rs.normal(0, 1, (cfg.context, D))
    tree = [[[etok], [epos]]]
    for _ in range(cfg.blocks):
        qkv = _fc_init(rs.uniform, D, 3 * D)
        fc0 = _fc_init(rs.uniform, D, cfg.mlp_ratio * D)
        fc1 = _fc_init(rs.uniform, cfg.mlp_ratio * D, D)
        fco = _fc_init(rs.uniform, D, D)
        tree.append([_ln_init(D), qkv, _ln_init(D), fc0, fc1, fco])
    tree.append(_ln_init(D))
    c0 = _fc_init(rs.uniform, D, D)
    c1 = _fc_init(rs.uniform, D, cfg.vocab)
    tree.append([c0, c1])
    return tree


def synthetic_vit_batch(cfg: ViTConfig, seed=1):
    """images rand(B,C,H,W) in [0,1), labels randint(0,classes,B) one-hot."""
    rs = np.random.RandomState(seed)
    images = rs.rand(cfg.batch, cfg.channels, cfg.height, cfg.width)
    labels = rs.randint(0, cfg.classes, cfg.batch)
    onehot = np.zeros((cfg.batch, cfg.classes))
    onehot[np.arange(cfg.batch), labels] = 1
    return images, labels, onehot


def synthetic_gpt_batch(cfg: GPTConfig, seed=1):
    rs = np.random.RandomState(seed)
    tokens = rs.randint(0, cfg.vocab, (cfg.batch, cfg.context))
    labels = rs.randint(0, cfg.vocab, (cfg.batch, cfg.context))
    return tokens, labels


def causal_mask(T, neg=-np.inf):
    """gpt/gpt_train_potry3000.py:84-91."""
    m = np.tril(np.ones((T, T)))
    out = np.zeros((T, T))
    out[m == 0] = neg
    return out


# ----------------------------------------------------------------------------- ViT
def vit_block_fwd(x, blk, heads, refstruct=False):
    """attention.py:20-55."""
    (g0, b0), (Wqkv, bqkv), (g1, b1), (W0, bb0), (W1, bb1) = blk
    u, xhat0, var0 = R.layernorm_fwd(x, g0, b0)
    qkv = R.linear_fwd(u, Wqkv, bqkv)
    if refstruct:
        O, P = R.attention_fwd_refstruct(qkv, heads)
    else:
        O, P, _ = R.attention_fwd(qkv, heads)
    h = x + O
    u1, xhat1, var1 = R.layernorm_fwd(h, g1, b1)
    pre = R.linear_fwd(u1, W0, bb0)
    act = R.gelu_fwd(pre)
    out = R.linear_fwd(act, W1, bb1) + h
    cache = dict(u=u, xhat0=xhat0, var0=var0, qkv=qkv, P=P, h=h, u1=u1, xhat1=xhat1, var1=var1,
                 pre=pre, act=act)
    return out, cache


def vit_block_bwd(delta, blk, c, heads, refstruct=False):
    """attention.py:57-89."""
    (g0, b0), (Wqkv, bqkv), (g1, b1), (W0, bb0), (W1, bb1) = blk
    d, dW1, db1 = R.linear_bwd(delta, c['act'], W1)
    d = R.gelu_bwd(d, c['pre'])
    d, dW0, db0 = R.linear_bwd(d, c['u1'], W0)
    delta0, dg1, dbt1 = R.layernorm_bwd(d, c['xhat1'], c['var1'], g1)
    delta0 = delta0 + delta
    if refstruct:
        dqkv = R.attention_bwd_refstruct(delta0, c['qkv'], c['P'], heads)
    else:
        dqkv = R.attention_bwd(delta0, c['qkv'], c['P'], heads)
    d, dWqkv, dbqkv = R.linear_bwd(dqkv, c['u'], Wqkv)
    dx, dg0, dbt0 = R.layernorm_bwd(d, c['xhat0'], c['var0'], g0)
    dx = dx + delta0
    grads = [[dg0, dbt0], [dWqkv, dbqkv], [dg1, dbt1], [dW0, db0], [dW1, db1]]
    return dx, grads, dict(dqkv=dqkv, delta0=delta0)


def vit_forward(params, images, cfg: ViTConfig, refstruct=False):
    """transformer_of_image.py:141-148 walked over PatchEmbed.py:22-43, Position_add.py:20-23,
    attention.py:20-55, net/layernorm.py, net/flatten.py, classify.py:20-28."""
    (gpe, bpe), (Wpe, bbpe) = params[0]
    acts = {}
    patches = R.patchify(images, cfg.n_patch)
    pe_lin = R.linear_fwd(patches, Wpe, bbpe)
    pe, xhat_pe, var_pe = R.layernorm_fwd(pe_lin, gpe, bpe)
    x = pe + R.posk_table(cfg.tokens, cfg.embed_dim)
    acts['patches'], acts['xhat_pe'], acts['var_pe'] = patches, xhat_pe, var_pe
    layer_out = [pe, x]
    caches = []
    for l in range(cfg.blocks):
        x, c = vit_block_fwd(x, params[2 + l], cfg.heads, refstruct)
        caches.append(c)
        layer_out.append(x)
    gf, bf = params[2 + cfg.blocks]
    z, xhat_f, var_f = R.layernorm_fwd(x, gf, bf)
    layer_out.append(z)
    zf = z.reshape(z.shape[0], -1)
    (Wc0, bc0), (Wc1, bc1) = params[3 + cfg.blocks]
    hpre = R.linear_fwd(zf, Wc0, bc0)
    hact = R.relu_fwd(hpre)
    logits = R.linear_fwd(hact, Wc1, bc1)
    layer_out.append(logits)
    acts.update(caches=caches, xhat_f=xhat_f, var_f=var_f, zf=zf, hpre=hpre, hact=hact,
                layer_out=layer_out)
    return logits, acts


def vit_backward(params, acts, dlogits, cfg: ViTConfig, refstruct=False, head_relu_mask=None):
    """transformer_of_image.py:156-159 (backward halves of the same layers).  head_relu_mask: test-only override of the
    head's ReLU mask (see gpt_block_bwd)."""
    (Wc0, bc0), (Wc1, bc1) = params[3 + cfg.blocks]
    d, dWc1, dbc1 = R.linear_bwd(dlogits, acts['hact'], Wc1)
    d = R.relu_bwd(d, acts['hpre']) if head_relu_mask is None else d * head_relu_mask.reshape(d.shape)
    d, dWc0, dbc0 = R.linear_bwd(d, acts['zf'], Wc0)
    deltas = [d]
    d = d.reshape(-1, cfg.tokens, cfg.embed_dim)
    gf, bf = params[2 + cfg.blocks]
    d, dgf, dbf = R.layernorm_bwd(d, acts['xhat_f'], acts['var_f'], gf)
    deltas.append(d)
    blk_grads = []
    for l in reversed(range(cfg.blocks)):
        d, g, _ = vit_block_bwd(d, params[2 + l], acts['caches'][l], cfg.heads, refstruct)
        blk_grads.insert(0, g)
        deltas.append(d)
    (gpe, bpe), (Wpe, bbpe) = params[0]
    d, dgpe, dbpe = R.layernorm_bwd(d, acts['xhat_pe'], acts['var_pe'], gpe)
    dpatch, dWpe, dbbpe = R.linear_bwd(d, acts['patches'], Wpe)
    deltas.append(dpatch)
    grads = [[[dgpe, dbpe], [dWpe, dbbpe]], []] + blk_grads + [[dgf, dbf], [[dWc0, dbc0], [dWc1, dbc1]]]
    return grads, deltas


def vit_step(params, images, onehot, lr, cfg: ViTConfig, refstruct=False, head_relu_mask=None):
    """One SGD training step (transformer_of_image.py:141-159). Returns a dict with loss, logits,
    probs, argmax, grads, deltas, activations and the updated parameter tree."""
    logits, acts = vit_forward(params, images, cfg, refstruct)
    loss, partial, probs = R.cross_entropy(logits, onehot)
    grads, deltas = vit_backward(params, acts, partial, cfg, refstruct, head_relu_mask)
    new_params = tree_map(lambda p, g: R.sgd_update(p, g, lr), params, grads)
    return dict(loss=loss, logits=logits, probs=probs, argmax=np.argmax(probs, axis=-1),
                partial=partial, grads=grads, deltas=deltas, acts=acts, params=new_params)


# ----------------------------------------------------------------------------- GPT
def gpt_block_fwd(x, blk, heads, mask, refstruct=False):
    """gpt/attdecoderblock.py:31-75 (MLP first, then masked MHA + fc_out)."""
    (g0, b0), (Wqkv, bqkv), (g1, b1), (W0, bb0), (W1, bb1), (Wo, bo) = blk
    u, xhat0, var0 = R.layernorm_fwd(x, g0, b0)
    pre = R.linear_fwd(u, W0, bb0)
    act = R.relu_fwd(pre)
    h = x + R.linear_fwd(act, W1, bb1)
    u1, xhat1, var1 = R.layernorm_fwd(h, g1, b1)
    qkv = R.linear_fwd(u1, Wqkv, bqkv)
    if refstruct:
        O, P = R.attention_fwd_refstruct(qkv, heads, mask)
    else:
        O, P, _ = R.attention_fwd(qkv, heads, mask)
    out = h + R.linear_fwd(O, Wo, bo)
    cache = dict(u=u, xhat0=xhat0, var0=var0, pre=pre, act=act, h=h, u1=u1, xhat1=xhat1, var1=var1,
                 qkv=qkv, P=P, O=O)
    return out, cache


def gpt_block_bwd(delta, blk, c, heads, refstruct=False, relu_mask=None):
    """gpt/attdecoderblock.py:77-111.  relu_mask (test-only): use this boolean mask instead of `pre > 0` in the ReLU
    backward — lets a parity test hold every gradient to 1e-4 although a handful of pre-activations sit within fp32
    rounding of zero, where the derivative of ReLU is discontinuous (the test bounds those positions separately)."""
    (g0, b0), (Wqkv, bqkv), (g1, b1), (W0, bb0), (W1, bb1), (Wo, bo) = blk
    d0, dWo, dbo = R.linear_bwd(delta, c['O'], Wo)
    if refstruct:
        dqkv = R.attention_bwd_refstruct(d0, c['qkv'], c['P'], heads)
    else:
        dqkv = R.attention_bwd(d0, c['qkv'], c['P'], heads)
    d, dWqkv, dbqkv = R.linear_bwd(dqkv, c['u1'], Wqkv)
    d, dg1, dbt1 = R.layernorm_bwd(d, c['xhat1'], c['var1'], g1)
    dh = d + delta
    d, dW1, db1 = R.linear_bwd(dh, c['act'], W1)
    d = R.relu_bwd(d, c['pre']) if relu_mask is None else d * relu_mask.reshape(d.shape)
    d, dW0, db0 = R.linear_bwd(d, c['u'], W0)
    dx, dg0, dbt0 = R.layernorm_bwd(d, c['xhat0'], c['var0'], g0)
    dx = dx + dh
    grads = [[dg0, dbt0], [dWqkv, dbqkv], [dg1, dbt1], [dW0, db0], [dW1, db1], [dWo, dbo]]
    return dx, grads


def gpt_forward(params, tokens, cfg: GPTConfig, mask, refstruct=False):
    """gpt/gpt_train_potry3000.py:256-262 over PatchEmbed.py:132-138, gpt/attdecoderblock.py,
    net/layernorm.py, gpt/classify_gpt.py:20-28 (relu=False, per-token head)."""
    (etok,), (epos,) = params[0]
    T = tokens.shape[1]
    x = R.embedding_fwd(tokens, etok) + epos[:T][None]
    layer_out = [x]
    caches = []
    for l in range(cfg.blocks):
        x, c = gpt_block_fwd(x, params[1 + l], cfg.heads, mask, refstruct)
        caches.append(c)
        layer_out.append(x)
    gf, bf = params[1 + cfg.blocks]
    z, xhat_f, var_f = R.layernorm_fwd(x, gf, bf)
    layer_out.append(z)
    (Wc0, bc0), (Wc1, bc1) = params[2 + cfg.blocks]
    h = R.linear_fwd(z, Wc0, bc0)
    logits = R.linear_fwd(h, Wc1, bc1)
    layer_out.append(logits)
    acts = dict(caches=caches, xhat_f=xhat_f, var_f=var_f, z=z, h=h, layer_out=layer_out)
    return logits, acts


def gpt_backward(params, acts, dlogits, tokens, cfg: GPTConfig, refstruct=False, relu_masks=None):
    (Wc0, bc0), (Wc1, bc1) = params[2 + cfg.blocks]
    d, dWc1, dbc1 = R.linear_bwd(dlogits, acts['h'], Wc1)
    d, dWc0, dbc0 = R.linear_bwd(d, acts['z'], Wc0)
    gf, bf = params[1 + cfg.blocks]
    d, dgf, dbf = R.layernorm_bwd(d, acts['xhat_f'], acts['var_f'], gf)
    blk_grads = []
    for l in reversed(range(cfg.blocks)):
        d, g = gpt_block_bwd(d, params[1 + l], acts['caches'][l], cfg.heads, refstruct,
                             None if relu_masks is None else relu_masks[l])
        blk_grads.insert(0, g)
    (etok,), (epos,) = params[0]
    detok = R.embedding_bwd(tokens, d, etok.shape[0])          # PatchEmbed.py:141
    depos = np.zeros_like(epos)
    depos[:tokens.shape[1]] = d.sum(axis=0)                    # PatchEmbed.py:142-143
    grads = [[[detok], [depos]]] + blk_grads + [[dgf, dbf], [[dWc0, dbc0], [dWc1, dbc1]]]
    return grads, d


def gpt_step(params, opt, tokens, labels, lr, cfg: GPTConfig, mask=None, refstruct=False, adam=True, relu_masks=None):
    """One training step of gpt/gpt_train_potry3000.py:253-288.  `opt` = dict(m=tree, v=tree, t=int)
    (Adam* state; every layer object's `t` advances in lock-step so one counter suffices)."""
    B, T = tokens.shape
    if mask is None:
        mask = causal_mask(T)
    logits, acts = gpt_forward(params, tokens, cfg, mask, refstruct)
    flat = logits.reshape(-1, cfg.vocab)
    onehot = np.zeros_like(flat)
    onehot[np.arange(flat.shape[0]), labels.reshape(-1)] = 1
    loss, partial, probs = R.cross_entropy(flat, onehot)
    grads, dx0 = gpt_backward(params, acts, partial.reshape(logits.shape), tokens, cfg, refstruct, relu_masks)
    if adam:
        t = opt['t']
        trip = tree_map(lambda p, g, m, v: R.adam_star_update(p, g, m, v, t, lr),
                        params, grads, opt['m'], opt['v'])
        # split the per-leaf (p, m, v) triples back into three trees
        new_params = _pick(trip, params, 0)
        new_opt = dict(m=_pick(trip, params, 1), v=_pick(trip, params, 2), t=t + 1)
    else:
        new_params = tree_map(lambda p, g: R.sgd_update(p, g, lr), params, grads)
        new_opt = opt
    return dict(loss=loss, logits=logits, probs=probs, argmax=np.argmax(probs, axis=-1),
                grads=grads, acts=acts, params=new_params, opt=new_opt, dx0=dx0)


def _pick(trip_tree, like, k):
    if isinstance(like, (list, tuple)):
        return [_pick(trip_tree[i], like[i], k) for i in range(len(like))]
    return trip_tree[k]


def adam_state(params):
    return dict(m=tree_zeros_like(params), v=tree_zeros_like(params), t=1)


def tree_unflatten(like, leaves):
    """Rebuild a tree shaped like `like` from a flat list of leaves (inverse of tree_leaves)."""
    it = iter(leaves)

    def rec(t):
        if isinstance(t, (list, tuple)):
            return [rec(x) for x in t]
        return next(it)
    return rec(like)
