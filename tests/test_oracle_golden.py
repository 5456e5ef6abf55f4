"""Pins oracle/ against vectors produced by the unmodified reference (oracle/gen_golden.py).
CPU only.  Bar: fp64 agreement to 1e-12 (max|a-b|/max|b|)."""
import numpy as np
import pytest
from conftest import load_golden, golden_tree, rel_err
from oracle import numpy_ref as R
from oracle import models as M

TOL = 1e-12


def test_linear(prims):
    p = prims
    assert rel_err(R.linear_fwd(p["lin_x"], p["lin_W"], p["lin_b"]), p["lin_y"]) < TOL
    dx, dW, db = R.linear_bwd(p["lin_dy"], p["lin_x"], p["lin_W"])
    assert rel_err(dx, p["lin_dx"]) < TOL
    assert rel_err(dW, p["lin_dW"]) < TOL
    assert rel_err(db, p["lin_db"]) < TOL


def test_optimisers(prims):
    p = prims
    Ws, bs = p["opt_W0"], p["opt_b0"]
    Wa, ba = Ws.copy(), bs.copy()
    mW = vW = np.zeros_like(Ws)
    mb = vb = np.zeros_like(bs)
    for t in range(3):
        Ws = R.sgd_update(Ws, p["opt_gW"][t], 0.01)
        bs = R.sgd_update(bs, p["opt_gb"][t], 0.01)
        Wa, mW, vW = R.adam_star_update(Wa, p["opt_gW"][t], mW, vW, t + 1, 0.01)
        ba, mb, vb = R.adam_star_update(ba, p["opt_gb"][t], mb, vb, t + 1, 0.01)
        assert rel_err(Ws, p["sgd_W_%d" % t]) < TOL and rel_err(bs, p["sgd_b_%d" % t]) < TOL
        assert rel_err(Wa, p["adam_W_%d" % t]) < TOL and rel_err(ba, p["adam_b_%d" % t]) < TOL


def test_layernorm(prims):
    p = prims
    y, xhat, var = R.layernorm_fwd(p["ln_x"], p["ln_g"], p["ln_b"])
    assert rel_err(y, p["ln_y"]) < TOL
    dx, dg, db = R.layernorm_bwd(p["ln_dy"], xhat, var, p["ln_g"])
    assert rel_err(dx, p["ln_dx"]) < TOL
    assert rel_err(dg, p["ln_dg"]) < TOL and rel_err(db, p["ln_db"]) < TOL


def test_softmax_gelu_relu(prims):
    p = prims
    assert rel_err(R.softmax_fwd(p["sm_x"]), p["sm_p"]) < TOL
    assert rel_err(R.softmax_bwd(p["sm_g"], p["sm_p"]), p["sm_dx"]) < 1e-11
    assert rel_err(R.softmax_bwd_refstruct(p["sm_g"], p["sm_p"]), p["sm_dx"]) < TOL
    assert rel_err(R.gelu_fwd(p["gelu_x"]), p["gelu_y"]) < TOL
    assert rel_err(R.gelu_bwd(p["gelu_g"], p["gelu_x"]), p["gelu_dx"]) < TOL
    assert np.array_equal(R.relu_fwd(p["relu_x"]), p["relu_y"])
    assert np.array_equal(R.relu_bwd(p["gelu_g"], p["relu_x"]), p["relu_dx"])


def test_cross_entropy(prims):
    p = prims
    loss, partial, sp = R.cross_entropy(p["ce_logits"], p["ce_onehot"])
    assert abs(loss - p["ce_loss"]) < 1e-14
    assert rel_err(partial, p["ce_partial"]) < TOL and rel_err(sp, p["ce_p"]) < TOL
    assert np.array_equal(np.argmax(sp, -1), np.argmax(p["ce_p"], -1))


def test_embedding_patch_pos(prims):
    p = prims
    assert np.array_equal(R.embedding_fwd(p["emb_idx"], p["emb_E"]), p["emb_y"])
    assert rel_err(R.embedding_bwd(p["emb_idx"], p["emb_g"], 9), p["emb_dE"]) < TOL
    assert np.array_equal(R.patchify(p["pe_img"], 3), p["pe_patches"])
    flat = p["pe_params"]
    g, b, W, bb = flat[:8], flat[8:16], flat[16:16 + 32 * 8].reshape(32, 8), flat[16 + 256:]
    y, _, _ = R.layernorm_fwd(R.linear_fwd(p["pe_patches"], W, bb), g, b)
    assert rel_err(y, p["pe_y"]) < TOL
    assert rel_err(R.posk_table(9, 8), p["posk"]) < TOL


def _blk_like(n):
    return [[0, 0]] * n


@pytest.mark.parametrize("refstruct", [False, True])
def test_vit_block(prims, refstruct):
    p = prims
    blk = golden_tree(p, "vb_p", _blk_like(5))
    blk = [[a.reshape(-1) if a.ndim == 3 else a for a in pair] for pair in blk]
    y, c = M.vit_block_fwd(p["vb_x"], blk, 2, refstruct)
    assert rel_err(y, p["vb_y"]) < TOL
    assert rel_err(c["qkv"], p["vb_qkv"]) < TOL
    assert rel_err(np.array(c["P"]), p["vb_P"]) < TOL
    dx, grads, extra = M.vit_block_bwd(p["vb_g"], blk, c, 2, refstruct)
    assert rel_err(dx, p["vb_dx"]) < 1e-11
    assert rel_err(extra["dqkv"], p["vb_dqkv"]) < 1e-11
    for a, b in zip(M.tree_leaves(grads), M.tree_leaves(golden_tree(p, "vb_grad", _blk_like(5)))):
        assert rel_err(a, b) < 1e-11


@pytest.mark.parametrize("refstruct", [False, True])
def test_gpt_block(prims, refstruct):
    p = prims
    blk = golden_tree(p, "gb_p", _blk_like(6))
    blk = [[a.reshape(-1) if a.ndim == 3 else a for a in pair] for pair in blk]
    mask = M.causal_mask(6)
    y, c = M.gpt_block_fwd(p["gb_x"], blk, 2, mask, refstruct)
    assert rel_err(y, p["gb_y"]) < TOL
    assert rel_err(np.array(c["P"]), p["gb_P"]) < TOL
    assert np.all(np.isfinite(y))
    dx, grads = M.gpt_block_bwd(p["gb_g"], blk, c, 2, refstruct)
    assert rel_err(dx, p["gb_dx"]) < 1e-11
    for a, b in zip(M.tree_leaves(grads), M.tree_leaves(golden_tree(p, "gb_grad", _blk_like(6)))):
        assert rel_err(a, b) < 1e-11


def _squeeze_tree(t):
    return M.tree_map(lambda a: a.reshape(-1) if (a.ndim == 3 and a.shape[0] == 1) else a, t)


def test_vit_tiny_two_steps():
    d = load_golden("vit_tiny.npz")
    cfg = M.ViTConfig(batch=3, embed_dim=16, heads=2, blocks=2, classes=10)
    like = M.vit_init(cfg, 0)
    params = golden_tree(d, "p0", like)
    for a, b in zip(M.tree_leaves(params), M.tree_leaves(like)):
        assert np.array_equal(a.reshape(-1), b.reshape(-1))      # initialiser parity
    out = M.vit_step(like, d["images"], d["onehot"], float(d["lr"]), cfg)
    assert abs(out["loss"] - d["loss_0"]) < 1e-13
    assert rel_err(out["logits"], d["logits_0"]) < TOL
    assert np.array_equal(out["argmax"], np.argmax(d["probs_0"], -1))
    # per-layer forward outputs: patchemb, pos, blocks, final norm, (flatten), logits
    lo = out["acts"]["layer_out"]
    gold_out = [d["out_%02d" % i] for i in range(2 + cfg.blocks + 3)]
    for i in range(2 + cfg.blocks + 1):
        assert rel_err(lo[i], gold_out[i]) < TOL, i
    # per-layer deltas (reverse order): classify, flatten, norm, blocks..., pos, patchemb
    gd = [d["delta_%02d" % i] for i in range(2 + cfg.blocks + 3)]
    assert rel_err(out["deltas"][0], gd[0]) < 1e-11
    assert rel_err(out["deltas"][1], gd[2]) < 1e-11
    for l in range(cfg.blocks):
        assert rel_err(out["deltas"][2 + l], gd[3 + l]) < 1e-11
    assert rel_err(out["deltas"][-1], gd[-1]) < 1e-11
    for a, b in zip(M.tree_leaves(out["grads"]), M.tree_leaves(golden_tree(d, "g0", like))):
        assert rel_err(a, b) < 1e-11
    p1 = _squeeze_tree(golden_tree(d, "p1", like))
    for a, b in zip(M.tree_leaves(out["params"]), M.tree_leaves(p1)):
        assert rel_err(a, b) < TOL
    out2 = M.vit_step(out["params"], d["images"], d["onehot"], float(d["lr"]), cfg, refstruct=True)
    assert abs(out2["loss"] - d["loss_1"]) < 1e-12
    p2 = _squeeze_tree(golden_tree(d, "p2", like))
    for a, b in zip(M.tree_leaves(out2["params"]), M.tree_leaves(p2)):
        assert rel_err(a, b) < 1e-11


def test_gpt_tiny_three_adam_steps():
    d = load_golden("gpt_tiny.npz")
    cfg = M.GPTConfig(batch=2, context=8, embed_dim=16, heads=2, blocks=2, vocab=11)
    params = M.gpt_init(cfg, 0)
    opt = M.adam_state(params)
    for s in range(3):
        out = M.gpt_step(params, opt, d["tokens"], d["labels"], float(d["lr"]), cfg)
        assert abs(out["loss"] - d["loss_%d" % s]) < 1e-12, s
        assert rel_err(out["logits"], d["logits_%d" % s]) < 1e-11
        if s == 0:
            for a, b in zip(M.tree_leaves(out["grads"]), M.tree_leaves(golden_tree(d, "g0", params))):
                assert rel_err(a, b) < 1e-11
        params, opt = out["params"], out["opt"]
        gold = _squeeze_tree(golden_tree(d, "p%d" % (s + 1), params))
        for a, b in zip(M.tree_leaves(params), M.tree_leaves(gold)):
            assert rel_err(a, b) < 1e-10, s


def test_vit_readme_trajectory_losses():
    """README-size ViT (B=100, D=96, 6x4 heads): loss/logits/argmax of two reference steps."""
    d = load_golden("vit_readme_traj.npz")
    cfg = M.VIT_README
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    for s in range(2):
        out = M.vit_step(params, images, onehot, float(d["lr"]), cfg)
        assert abs(out["loss"] - d["loss_%d" % s]) < 1e-11
        assert rel_err(out["logits"], d["logits_%d" % s]) < 1e-10
        assert np.array_equal(out["argmax"], d["argmax_%d" % s])
        if s == 0:
            gl = M.tree_leaves(out["grads"])
            assert np.allclose([np.abs(g).max() for g in gl], d["grad_absmax"], rtol=1e-9)
            assert rel_err(gl[4], d["grad_qkv0"]) < 1e-6      # stored as fp32
        params = out["params"]


CONV_CASES = [(4, 0), (1, 1), (2, 1), ((2, 1), (0, 1)), (3, 0), (2, 0)]      # (stride, padding) of oracle/gen_golden.py:CONV_CASES


@pytest.mark.parametrize("i", range(len(CONV_CASES)))
def test_convolution(i):
    """net/Convolution.py:104-131 forward, :163-225 backward (accumulating), :236-239 update."""
    g = load_golden("conv.npz")
    s, p = CONV_CASES[i]
    x, W, b = g["c%d_x" % i], g["c%d_W" % i], g["c%d_b" % i]
    assert rel_err(R.conv_fwd(x, W, b, s, p), g["c%d_y" % i]) < TOL
    dx, dW, db = R.conv_bwd(g["c%d_g" % i], x, W, s, p)
    assert rel_err(dx, g["c%d_dx" % i]) < TOL and rel_err(dW, g["c%d_dW" % i]) < TOL and rel_err(db, g["c%d_db" % i]) < TOL
    assert rel_err(2 * dW, g["c%d_dW2" % i]) < TOL
    assert rel_err(R.sgd_update(W, 2 * dW, 0.05), g["c%d_W1" % i]) < TOL
    assert rel_err(R.sgd_update(b, 2 * db, 0.05), g["c%d_b1" % i]) < TOL


def test_patch_embed_convolution():
    """PatchEmbed.py:73-122: conv with kernel = stride = patch, NCHW -> (n, patches, embed); forward hands back the tensor
    BEFORE the LayerNorm (:90-93 returns `out`), backward still runs the norm's backward on the incoming gradient."""
    g = load_golden("conv.npz")
    x, W, b = g["pe_x"], g["pe_W"], g["pe_b"]
    y = R.conv_fwd(x, W, b, 4, 0).transpose(0, 2, 3, 1).reshape(3, -1, 12)
    assert rel_err(y, g["pe_y"]) < TOL
    _, xhat, var = R.layernorm_fwd(y, np.ones(12), np.zeros(12))
    d, dgam, dbet = R.layernorm_bwd(g["pe_g"], xhat, var, np.ones(12))
    assert rel_err(dgam, g["pe_dgamma"]) < TOL and rel_err(dbet, g["pe_dbeta"]) < TOL
    dx, dW, db = R.conv_bwd(d.reshape(3, 2, 2, 12).transpose(0, 3, 1, 2), x, W, 4, 0)
    assert rel_err(dx, g["pe_dx"]) < TOL and rel_err(dW, g["pe_dW"]) < TOL and rel_err(db, g["pe_db"]) < TOL


def test_remaining_activations_and_losses():
    """net/activation.py:25-73 (sigmoid, tanh, SiLU), net/loss.py:53-57 (BCE on logits), :68-71 (MSE)."""
    g = load_golden("conv.npz")
    x, dy = g["act_x"], g["act_dy"]
    assert rel_err(R.sigmoid_fwd(x), g["sigmoid_y"]) < TOL and rel_err(R.sigmoid_bwd(dy, g["sigmoid_y"]), g["sigmoid_dx"]) < TOL
    assert rel_err(R.tanh_fwd(x), g["tanh_y"]) < TOL and rel_err(R.tanh_bwd(dy, g["tanh_y"]), g["tanh_dx"]) < TOL
    assert rel_err(R.silu_fwd(x), g["silu_y"]) < TOL and rel_err(R.silu_bwd(dy, x), g["silu_dx"]) < TOL
    loss, partial, s = R.bce_logit_loss(x, g["bce_label"])
    assert abs(loss - g["bce_loss"]) < TOL * abs(g["bce_loss"]) and rel_err(partial, g["bce_partial"]) < TOL
    assert rel_err(s, g["bce_sigmoid"]) < TOL
    loss, partial = R.mse_loss(x, g["mse_label"])
    assert abs(loss - g["mse_loss"]) < TOL * abs(g["mse_loss"]) and rel_err(partial, g["mse_partial"]) < TOL
