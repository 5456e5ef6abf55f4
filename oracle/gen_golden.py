"""Generate tests/golden/*.npz by running the UNMODIFIED reference (imported from
/root/reference) on seeded inputs.  TEST INFRASTRUCTURE.  Run in the build container only:

    python -m oracle.gen_golden            # ~1 min (the README-size ViT trajectory dominates)

The GPU box has no /root/reference; tests there read the committed .npz files.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NT_REFERENCE", "/root/reference")
sys.path.insert(0, os.path.join(HERE, "_mpl_stub"))
sys.path.insert(1, REF)
sys.path.insert(2, os.path.join(REF, "gpt"))

from net.fullconnect import fclayer                     # noqa: E402
from net.layernorm import layer_norm                    # noqa: E402
from net.activation import ReLU, Softmax, GELU          # noqa: E402
from net.loss import cross_entropy_loss                 # noqa: E402
from net.flatten import flatten_layer                   # noqa: E402
from net.Embedding import Embedding_layer               # noqa: E402
from attention import attention_layer                   # noqa: E402
from classify import classify_layer                     # noqa: E402
from PatchEmbed import PatchEmbed_flatten, Position_Embedding   # noqa: E402
from Position_add import Position_learnable             # noqa: E402
from gpt.attdecoderblock import attdecoderblock_layer   # noqa: E402
from gpt.classify_gpt import classify_layer as classify_layer_gpt  # noqa: E402

from oracle import models as M                          # noqa: E402

if not hasattr(np.lib, "pad"):          # net/Convolution.py:114,201 call np.lib.pad, which NumPy 2 dropped: harness shim
    np.lib.pad = np.pad
from net.Convolution import convolution_layer           # noqa: E402
from PatchEmbed import PatchEmbed_convolution           # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def leaves(tree):
    return [np.array(a, copy=True) for a in M.tree_leaves(tree)]


def put_tree(d, prefix, tree):
    for i, a in enumerate(leaves(tree)):
        d["%s_%03d" % (prefix, i)] = a


def save_walk(layers):
    return [l.save_model() for l in layers if 'save_model' in dir(l) and 'restore_model' in dir(l)]


# ------------------------------------------------------------------ primitives
def gen_prims():
    d = {}
    np.random.seed(11)
    # Linear, 3-D input, fullconnect.py:99-127
    x = np.random.randn(3, 5, 12)
    fc = fclayer(12, 7, True)
    d["lin_x"], d["lin_W"], d["lin_b"] = x, fc.params.copy(), fc.bias_params.copy()
    d["lin_y"] = fc.forward(x)
    dy = np.random.randn(3, 5, 7)
    d["lin_dy"] = dy
    d["lin_dx"] = fc.backward(dy)
    d["lin_dW"], d["lin_db"] = fc.params_delta.copy(), fc.bias_delta.copy()
    # SGD + Adam* three steps, fullconnect.py:134-153
    fs = fclayer(6, 4, True)
    fa = fclayer(6, 4, True, adam=True)
    d["opt_W0"], d["opt_b0"] = fs.params.copy(), fs.bias_params.copy()
    fa.params, fa.bias_params = fs.params.copy(), fs.bias_params.copy()
    gs = np.random.randn(3, 6, 4)
    gb = np.random.randn(3, 4)
    d["opt_gW"], d["opt_gb"] = gs, gb
    for t in range(3):
        for f in (fs, fa):
            f.params_delta[...] = gs[t]
            f.bias_delta[...] = gb[t]
            f.update(0.01)
            f.setzero()
        d["sgd_W_%d" % t], d["sgd_b_%d" % t] = fs.params.copy(), fs.bias_params.copy()
        d["adam_W_%d" % t], d["adam_b_%d" % t] = fa.params.copy(), fa.bias_params.copy()
    # LayerNorm, layernorm.py:88-135
    x = np.random.randn(4, 6, 10) * 2 + 0.5
    ln = layer_norm(10)
    ln.gamma = np.random.randn(10)
    ln.beta = np.random.randn(10)
    d["ln_x"], d["ln_g"], d["ln_b"] = x, ln.gamma.copy(), ln.beta.copy()
    d["ln_y"] = ln.forward(x)
    dy = np.random.randn(4, 6, 10)
    d["ln_dy"] = dy
    d["ln_dx"] = ln.backward(dy)
    d["ln_dg"], d["ln_db"] = ln.gamma_delta.copy(), ln.beta_delta.copy()
    # Softmax fwd / Jacobian bwd, activation.py:76-129
    sm = Softmax()
    x = np.random.randn(9, 13) * 3
    p = sm.forward(x, axis=-1)
    g = np.random.randn(9, 13)
    d["sm_x"], d["sm_p"], d["sm_g"] = x, p.copy(), g
    d["sm_dx"] = sm.backward(g, p)
    # GELU / ReLU, activation.py:139-148, 11-17
    ge = GELU()
    x = np.random.randn(5, 17) * 2
    d["gelu_x"] = x
    d["gelu_y"] = ge.forward(x)
    g = np.random.randn(5, 17)
    d["gelu_g"] = g
    d["gelu_dx"] = ge.backward(g)
    re = ReLU()
    x = np.random.randn(5, 17)
    d["relu_x"] = x.copy()
    d["relu_y"] = re.forward(x.copy())
    d["relu_dx"] = re.backward(g)
    # cross entropy, loss.py:46-51
    logits = np.random.randn(8, 10) * 4
    lab = np.random.randint(0, 10, 8)
    oh = np.zeros((8, 10))
    oh[np.arange(8), lab] = 1
    loss, partial, sp = cross_entropy_loss(logits, oh)
    d["ce_logits"], d["ce_onehot"], d["ce_loss"], d["ce_partial"], d["ce_p"] = logits, oh, loss, partial, sp
    # Embedding, Embedding.py:50-70
    em = Embedding_layer(9, 6)
    idx = np.random.randint(0, 9, (3, 7))
    d["emb_E"], d["emb_idx"] = em.params.copy(), idx
    d["emb_y"] = em.forward(idx)
    g = np.random.randn(3, 7, 6)
    d["emb_g"] = g
    d["emb_dE"] = em.backward(g).copy()
    # patchify + Linear + LN, PatchEmbed.py:22-49 ; posk, Position_add.py:9-23
    img = np.random.rand(2, 2, 12, 12)
    pe = PatchEmbed_flatten(8, img.shape, 3, True)
    d["pe_img"] = img
    d["pe_params"] = np.concatenate([a.reshape(-1) for a in leaves(pe.save_model())])
    d["pe_y"] = pe.forward(img)
    d["pe_patches"] = pe.inputs.copy()
    pl = Position_learnable(3, 8, fixed=1)
    d["posk"] = pl.posk.copy()
    # ViT block, attention.py
    x = np.random.randn(2, 5, 8)
    at = attention_layer(8, 2)
    put_tree(d, "vb_p", at.save_model())
    d["vb_x"] = x
    d["vb_y"] = at.forward(x)
    d["vb_qkv"] = at.qkv.reshape(2, 5, -1).copy()
    d["vb_P"] = np.array(at.atg__)
    g = np.random.randn(2, 5, 8)
    d["vb_g"] = g
    d["vb_dx"] = at.backward(g.copy())
    d["vb_dqkv"] = at.delta_qkv.reshape(2, 5, -1).copy()
    put_tree(d, "vb_grad", [[at.norm.gamma_delta.reshape(-1), at.norm.beta_delta.reshape(-1)],
                            [at.qkvfc.params_delta, at.qkvfc.bias_delta],
                            [at.norm1.gamma_delta.reshape(-1), at.norm1.beta_delta.reshape(-1)],
                            [at.fc0.params_delta, at.fc0.bias_delta],
                            [at.fc1.params_delta, at.fc1.bias_delta]])
    # GPT block with -inf causal mask, gpt/attdecoderblock.py
    x = np.random.randn(2, 6, 8)
    mask = M.causal_mask(6)
    gb_ = attdecoderblock_layer(8, 2)
    put_tree(d, "gb_p", gb_.save_model())
    d["gb_x"] = x
    d["gb_y"] = gb_.forward(x, mask)
    d["gb_P"] = np.array(gb_.atg__)
    g = np.random.randn(2, 6, 8)
    d["gb_g"] = g
    d["gb_dx"] = gb_.backward(g.copy())
    put_tree(d, "gb_grad", [[gb_.norm.gamma_delta.reshape(-1), gb_.norm.beta_delta.reshape(-1)],
                            [gb_.qkvfc.params_delta, gb_.qkvfc.bias_delta],
                            [gb_.norm1.gamma_delta.reshape(-1), gb_.norm1.beta_delta.reshape(-1)],
                            [gb_.fc0.params_delta, gb_.fc0.bias_delta],
                            [gb_.fc1.params_delta, gb_.fc1.bias_delta],
                            [gb_.fc_out.params_delta, gb_.fc_out.bias_delta]])
    np.savez_compressed(os.path.join(OUT, "prims.npz"), **d)
    print("prims.npz", len(d), "arrays")


# ------------------------------------------------------------------ ViT model
def build_vit(cfg):
    """transformer_of_image.py:56-77 (patch_convolu=0, patchnorm, fixed=1, cls_token=0)."""
    shape = (cfg.batch, cfg.channels, cfg.height, cfg.width)
    layers = [PatchEmbed_flatten(cfg.embed_dim, shape, cfg.n_patch, patchnorm=True),
              Position_learnable(cfg.n_patch, cfg.embed_dim, fixed=1)]
    layers += [attention_layer(cfg.embed_dim, cfg.heads) for _ in range(cfg.blocks)]
    layers += [layer_norm(cfg.embed_dim), flatten_layer(),
               classify_layer(cfg.embed_dim, cfg.batch, cfg.n_patch, cfg.classes, 0)]
    return layers


def vit_grad_tree(layers, cfg):
    fl = lambda a: np.array(a, copy=True).reshape(-1)
    pe = layers[0]
    tree = [[[fl(pe.norm.gamma_delta), fl(pe.norm.beta_delta)],
             [pe.fullconnect.params_delta.copy(), pe.fullconnect.bias_delta.copy()]], []]
    for at in layers[2:2 + cfg.blocks]:
        tree.append([[fl(at.norm.gamma_delta), fl(at.norm.beta_delta)],
                     [at.qkvfc.params_delta.copy(), at.qkvfc.bias_delta.copy()],
                     [fl(at.norm1.gamma_delta), fl(at.norm1.beta_delta)],
                     [at.fc0.params_delta.copy(), at.fc0.bias_delta.copy()],
                     [at.fc1.params_delta.copy(), at.fc1.bias_delta.copy()]])
    ln = layers[2 + cfg.blocks]
    cl = layers[-1]
    tree.append([fl(ln.gamma_delta), fl(ln.beta_delta)])
    tree.append([[cl.fc0.params_delta.copy(), cl.fc0.bias_delta.copy()],
                 [cl.fc1.params_delta.copy(), cl.fc1.bias_delta.copy()]])
    return tree


def ref_vit_step(layers, images, onehot, lr, cfg, record=None):
    x = images
    outs = []
    for l in layers:
        x = l.forward(x)
        outs.append(np.array(x, copy=True))
    loss, delta, predict = cross_entropy_loss(x, onehot)
    logits = np.array(x, copy=True)
    deltas = []
    for l in reversed(layers):
        delta = l.backward(delta)
        deltas.append(np.array(delta, copy=True))
    grads = vit_grad_tree(layers, cfg)
    for l in layers:
        l.update(lr)
        l.setzero()
    if record is not None:
        record.update(outs=outs, deltas=deltas)
    return loss, logits, predict, grads


def gen_vit_tiny():
    cfg = M.ViTConfig(batch=3, embed_dim=16, heads=2, blocks=2, classes=10)
    np.random.seed(0)
    layers = build_vit(cfg)
    p0 = save_walk(layers)
    mine = M.vit_init(cfg, seed=0)
    for a, b in zip(leaves(p0), leaves(mine)):
        assert np.array_equal(a.reshape(-1), b.reshape(-1)), "oracle vit_init diverges from the reference constructors"
    images, labels, onehot = M.synthetic_vit_batch(cfg, seed=1)
    d = dict(images=images, labels=labels, onehot=onehot, lr=np.float64(0.05))
    put_tree(d, "p0", p0)
    rec = {}
    loss, logits, predict, grads = ref_vit_step(layers, images, onehot, 0.05, cfg, rec)
    d["loss_0"], d["logits_0"], d["probs_0"] = loss, logits, predict
    put_tree(d, "g0", grads)
    # per-layer outputs: patchemb, pos, blocks..., norm, flatten, logits ; deltas in reverse order
    for i, o in enumerate(rec["outs"]):
        d["out_%02d" % i] = o
    for i, o in enumerate(rec["deltas"]):
        d["delta_%02d" % i] = o
    put_tree(d, "p1", save_walk(layers))
    loss, logits, predict, grads = ref_vit_step(layers, images, onehot, 0.05, cfg)
    d["loss_1"], d["logits_1"] = loss, logits
    put_tree(d, "p2", save_walk(layers))
    np.savez_compressed(os.path.join(OUT, "vit_tiny.npz"), **d)
    print("vit_tiny.npz loss", d["loss_0"], d["loss_1"])


def gen_vit_readme():
    cfg = M.VIT_README
    np.random.seed(0)
    layers = build_vit(cfg)
    mine = M.vit_init(cfg, seed=0)
    for a, b in zip(leaves(save_walk(layers)), leaves(mine)):
        assert np.array_equal(a.reshape(-1), b.reshape(-1))
    images, labels, onehot = M.synthetic_vit_batch(cfg, seed=1)
    d = dict(lr=np.float64(1e-3))
    for s in range(2):
        loss, logits, predict, grads = ref_vit_step(layers, images, onehot, 1e-3, cfg)
        d["loss_%d" % s], d["logits_%d" % s] = loss, logits
        d["argmax_%d" % s] = np.argmax(predict, axis=-1)
        if s == 0:
            gl = leaves(grads)
            d["grad_absmax"] = np.array([np.abs(g).max() for g in gl])
            d["grad_sum"] = np.array([g.sum() for g in gl])
            d["grad_qkv0"] = gl[4].astype(np.float32)      # block-0 Wqkv grad (96,288)
            d["grad_wc1"] = gl[-2].astype(np.float32)
        print("vit README step", s, "loss", loss)
    np.savez_compressed(os.path.join(OUT, "vit_readme_traj.npz"), **d)


# ------------------------------------------------------------------ GPT model
def build_gpt(cfg, adam=True):
    """gpt/gpt_train_potry3000.py:176-207 with float32=False (fp64 goldens)."""
    layers = [Position_Embedding(cfg.context, cfg.vocab, cfg.embed_dim, adam=adam)]
    layers += [attdecoderblock_layer(cfg.embed_dim, cfg.heads, adam=adam) for _ in range(cfg.blocks)]
    layers += [layer_norm(cfg.embed_dim, adam=adam),
               classify_layer_gpt(cfg.embed_dim, cfg.batch, 1, cfg.vocab, cls_token=False, adam=adam, relu=False)]
    return layers


def gpt_grad_tree(layers, cfg):
    fl = lambda a: np.array(a, copy=True).reshape(-1)
    pe = layers[0]
    tree = [[[pe.text_embedding.delta.copy()], [pe.pos_embedding.delta.copy()]]]
    for at in layers[1:1 + cfg.blocks]:
        tree.append([[fl(at.norm.gamma_delta), fl(at.norm.beta_delta)],
                     [at.qkvfc.params_delta.copy(), at.qkvfc.bias_delta.copy()],
                     [fl(at.norm1.gamma_delta), fl(at.norm1.beta_delta)],
                     [at.fc0.params_delta.copy(), at.fc0.bias_delta.copy()],
                     [at.fc1.params_delta.copy(), at.fc1.bias_delta.copy()],
                     [at.fc_out.params_delta.copy(), at.fc_out.bias_delta.copy()]])
    ln, cl = layers[-2], layers[-1]
    tree.append([fl(ln.gamma_delta), fl(ln.beta_delta)])
    tree.append([[cl.fc0.params_delta.copy(), cl.fc0.bias_delta.copy()],
                 [cl.fc1.params_delta.copy(), cl.fc1.bias_delta.copy()]])
    return tree


def gen_gpt_tiny():
    cfg = M.GPTConfig(batch=2, context=8, embed_dim=16, heads=2, blocks=2, vocab=11)
    np.random.seed(0)
    layers = build_gpt(cfg)
    p0 = [l.save_model() for l in layers]
    # Embedding.save_model down-casts to fp32 (Embedding.py:93): take the fp64 tables directly
    p0[0] = [[layers[0].text_embedding.params.copy()], [layers[0].pos_embedding.params.copy()]]
    mine = M.gpt_init(cfg, seed=0)
    for a, b in zip(leaves(p0), leaves(mine)):
        assert np.array_equal(a.reshape(-1), b.reshape(-1)), "oracle gpt_init diverges from the reference constructors"
    tokens, labels = M.synthetic_gpt_batch(cfg, seed=1)
    mask = M.causal_mask(cfg.context)
    d = dict(tokens=tokens, labels=labels, lr=np.float64(3e-3))
    put_tree(d, "p0", p0)
    for s in range(3):
        x = tokens
        for l in layers:
            x = l.forward(x, mask) if isinstance(l, attdecoderblock_layer) else l.forward(x)
        ishape = x.shape
        flat = np.reshape(x, (-1, cfg.vocab))
        oh = np.zeros_like(flat)
        oh[np.arange(len(flat)), labels.reshape(-1)] = 1
        loss, delta, predict = cross_entropy_loss(flat, oh)
        d["loss_%d" % s] = loss
        d["logits_%d" % s] = np.array(x, copy=True)
        delta = np.reshape(delta, ishape)
        for l in reversed(layers):
            delta = l.backward(delta)
        if s == 0:
            put_tree(d, "g0", gpt_grad_tree(layers, cfg))
        for l in layers:
            l.update(3e-3)
            l.setzero()
        ps = [l.save_model() for l in layers]
        ps[0] = [[layers[0].text_embedding.params.copy()], [layers[0].pos_embedding.params.copy()]]
        put_tree(d, "p%d" % (s + 1), ps)
    np.savez_compressed(os.path.join(OUT, "gpt_tiny.npz"), **d)
    print("gpt_tiny.npz losses", d["loss_0"], d["loss_1"], d["loss_2"])


# ------------------------------------------------------------------ convolution
CONV_CASES = [  # n, c, h, w, oc, kernel, stride, padding
    (2, 1, 8, 8, 5, 4, 4, 0), (2, 2, 6, 7, 3, 3, 1, 1), (1, 2, 7, 7, 3, 3, 2, 1), (2, 3, 9, 8, 4, (2, 3), (2, 1), (0, 1)),
    (1, 1, 10, 10, 2, 3, 3, 0), (1, 2, 8, 9, 3, 3, 2, 0)]


def gen_conv():
    """net/Convolution.py:37-248 forward / backward and PatchEmbed.py:73-122 (conv patch embedding, incl. its forward
    returning the PRE-LayerNorm tensor while backward still goes through the norm)."""
    d = {}
    for i, (n, c, h, w, oc, k, s, p) in enumerate(CONV_CASES):
        np.random.seed(100 + i)
        x = np.random.randn(n, c, h, w)
        conv = convolution_layer(c, oc, k, stride=s, padding=p)
        d["c%d_x" % i], d["c%d_W" % i], d["c%d_b" % i] = x, conv.params.copy(), conv.bias_params.copy()
        y = conv.forward(x)
        g = np.random.randn(*y.shape)
        d["c%d_y" % i], d["c%d_g" % i] = y, g
        d["c%d_dx" % i] = conv.backward(g)
        d["c%d_dW" % i], d["c%d_db" % i] = conv.params_delta.copy(), conv.bias_delta.copy()
        conv.backward(g)                                  # accumulate-until-setzero
        d["c%d_dW2" % i] = conv.params_delta.copy()
        conv.update(0.05)
        d["c%d_W1" % i], d["c%d_b1" % i] = conv.params.copy(), conv.bias_params.copy()
    np.random.seed(7)
    pe = PatchEmbed_convolution(12, (3, 2, 8, 8), 2, patchnorm=True)
    x = np.random.randn(3, 2, 8, 8)
    d["pe_x"], d["pe_W"], d["pe_b"] = x, pe.convolution.params.copy(), pe.convolution.bias_params.copy()
    d["pe_y"] = pe.forward(x)
    g = np.random.randn(3, 4, 12)
    d["pe_g"] = g
    d["pe_dx"] = pe.backward(g)
    d["pe_dW"], d["pe_db"] = pe.convolution.params_delta.copy(), pe.convolution.bias_delta.copy()
    d["pe_dgamma"], d["pe_dbeta"] = np.array(pe.norm.gamma_delta).reshape(-1), np.array(pe.norm.beta_delta).reshape(-1)
    # net/activation.py:25-73 sigmoid / tanh / SiLU; net/loss.py:53-57 BCE on logits, :68-71 MSE
    from net.activation import sigmoid, tanh, SiLU
    from net.loss import binary_cross_entropy_logit_loss, mean_square_loss
    np.random.seed(21)
    x, dy = np.random.randn(4, 9) * 2, np.random.randn(4, 9)
    d["act_x"], d["act_dy"] = x, dy
    for name, cls in (("sigmoid", sigmoid), ("tanh", tanh), ("silu", SiLU)):
        a = cls()
        d[name + "_y"] = a.forward(x.copy())
        d[name + "_dx"] = a.backward(dy)
    lab = (np.random.rand(4, 9) > 0.5).astype(np.float64)
    d["bce_label"] = lab
    d["bce_loss"], d["bce_partial"], d["bce_sigmoid"] = binary_cross_entropy_logit_loss(x, lab)
    tgt = np.random.randn(4, 9)
    d["mse_label"] = tgt
    d["mse_loss"], d["mse_partial"] = mean_square_loss(x, tgt)
    np.savez_compressed(os.path.join(OUT, "conv.npz"), **d)


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if "--only-conv" in sys.argv:
        gen_conv()
        sys.exit(0)
    gen_conv()
    gen_prims()
    gen_vit_tiny()
    gen_gpt_tiny()
    if "--skip-readme" not in sys.argv:
        gen_vit_readme()
