"""GPU parity, op level: the drop-in layer classes (through the C ABI) against the oracle and the
reference-generated golden vectors.  Tolerance for the fp32 path: max|a-b|/max|b| <= 1e-4
(BASELINE.json north_star); index results (argmax) bit-exact."""
import numpy as np
import pytest
from conftest import load_golden, golden_tree, rel_err
from oracle import numpy_ref as R
from oracle import models as M

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    return nt


def test_library_loaded_is_in_tree(nt):
    import os
    assert os.path.exists(nt.LIBPATH)
    maps = open("/proc/self/maps").read()
    assert "libntb200.so" in maps


def test_fclayer_golden(nt, prims):
    from net.fullconnect import fclayer
    p = prims
    fc = fclayer(12, 7, True, params=p["lin_W"], bias_params=p["lin_b"])
    y = fc.forward(p["lin_x"])
    assert y.shape == (3, 5, 7)
    assert rel_err(y, p["lin_y"]) < TOL
    dx = fc.backward(p["lin_dy"])
    assert rel_err(dx, p["lin_dx"]) < TOL
    assert rel_err(fc.params_delta, p["lin_dW"]) < TOL
    assert rel_err(fc.bias_delta, p["lin_db"]) < TOL
    # gradients accumulate until setzero (fullconnect.py:121-125)
    fc.backward(p["lin_dy"])
    assert rel_err(fc.params_delta, 2 * p["lin_dW"]) < TOL
    fc.setzero()
    assert not np.any(np.asarray(fc.params_delta))
    sm = fc.save_model()
    assert sm[0].dtype == np.float64 and sm[0].shape == (12, 7) and sm[1].shape == (7,)


@pytest.mark.parametrize("M_,K,N", [(1, 1, 1), (3, 5, 2), (49, 16, 96), (130, 96, 288), (100, 4704, 96), (257, 192, 10),
                                    (64, 65, 63)])
def test_linear_shapes_vs_oracle(nt, M_, K, N):
    from net.fullconnect import fclayer
    rs = np.random.RandomState(M_ * 7 + K)
    x, W, b = rs.randn(M_, K), rs.randn(K, N) / np.sqrt(K), rs.randn(N)
    dy = rs.randn(M_, N)
    fc = fclayer(K, N, True, params=W, bias_params=b)
    assert rel_err(fc.forward(x), R.linear_fwd(x, W, b)) < TOL
    dx = fc.backward(dy)
    rdx, rdW, rdb = R.linear_bwd(dy, x, W)
    assert rel_err(dx, rdx) < TOL and rel_err(fc.params_delta, rdW) < TOL and rel_err(fc.bias_delta, rdb) < TOL


def test_optimisers_golden(nt, prims):
    from net.fullconnect import fclayer
    p = prims
    fs = fclayer(6, 4, True, params=p["opt_W0"], bias_params=p["opt_b0"])
    fa = fclayer(6, 4, True, params=p["opt_W0"], bias_params=p["opt_b0"], adam=True)
    for t in range(3):
        for f in (fs, fa):
            f.params_delta.set(p["opt_gW"][t])
            f.bias_delta.set(p["opt_gb"][t])
            f.update(0.01)
            f.setzero()
        assert rel_err(fs.params, p["sgd_W_%d" % t]) < 1e-6 and rel_err(fs.bias_params, p["sgd_b_%d" % t]) < 1e-6
        assert rel_err(fa.params, p["adam_W_%d" % t]) < 1e-5 and rel_err(fa.bias_params, p["adam_b_%d" % t]) < 1e-5
    assert fa.t == 4


def test_layernorm_golden(nt, prims):
    from net.layernorm import layer_norm
    p = prims
    ln = layer_norm(10, gamma=p["ln_g"], beta=p["ln_b"])
    assert rel_err(ln.forward(p["ln_x"]), p["ln_y"]) < TOL
    assert rel_err(ln.backward(p["ln_dy"]), p["ln_dx"]) < TOL
    assert rel_err(ln.gamma_delta, p["ln_dg"]) < TOL and rel_err(ln.beta_delta, p["ln_db"]) < TOL
    g, b = ln.save_model()
    assert g.shape == (1, 1, 10) and b.shape == (1, 1, 10)     # rank-expanded like layernorm.py:105-108


@pytest.mark.parametrize("rows,D", [(1, 1), (7, 96), (4900, 96), (33, 384), (5, 1000)])
def test_layernorm_shapes(nt, rows, D):
    from net.layernorm import layer_norm
    rs = np.random.RandomState(rows + D)
    x, g, b, dy = rs.randn(rows, D) * 3 + 1, rs.randn(D), rs.randn(D), rs.randn(rows, D)
    ln = layer_norm(D, gamma=g, beta=b)
    y, xhat, var = R.layernorm_fwd(x, g, b)
    assert rel_err(ln.forward(x), y) < TOL
    dx, dg, db = R.layernorm_bwd(dy, xhat, var, g)
    assert rel_err(ln.backward(dy), dx) < 2e-4
    assert rel_err(ln.gamma_delta, dg) < TOL and rel_err(ln.beta_delta, db) < TOL


def test_activations_golden(nt, prims):
    from net.activation import ReLU, GELU, Softmax
    p = prims
    sm = Softmax()
    out = sm.forward(p["sm_x"], axis=-1)
    assert rel_err(out, p["sm_p"]) < TOL
    assert rel_err(sm.backward(p["sm_g"], p["sm_p"]), p["sm_dx"]) < TOL
    ge = GELU()
    assert rel_err(ge.forward(p["gelu_x"]), p["gelu_y"]) < TOL
    assert rel_err(ge.backward(p["gelu_g"]), p["gelu_dx"]) < TOL
    re = ReLU()
    x = p["relu_x"].copy()
    y = re.forward(x)
    assert rel_err(y, p["relu_y"]) < 1e-7
    assert np.array_equal(x, p["relu_y"])                       # in-place clamp of the caller's array (:13)
    assert rel_err(re.backward(p["gelu_g"]), p["relu_dx"]) < 1e-7
    for cls in (ReLU, GELU, Softmax):
        assert not hasattr(cls, "save_model")                   # checkpoint walkers must skip them


def test_cross_entropy_golden_and_argmax(nt, prims):
    from net.loss import cross_entropy_loss
    p = prims
    loss, partial, sp = cross_entropy_loss(p["ce_logits"], p["ce_onehot"])
    assert abs(float(loss) - p["ce_loss"]) < 1e-5 * abs(p["ce_loss"])
    assert rel_err(partial, p["ce_partial"]) < TOL and rel_err(sp, p["ce_p"]) < TOL
    assert np.array_equal(np.argmax(sp, axis=-1), np.argmax(p["ce_p"], axis=-1))
    assert "{:.6f}".format(loss) == "{:.6f}".format(p["ce_loss"])
    # soft (non one-hot) labels take the dense path
    soft = p["ce_onehot"] * 0.9 + 0.01
    l2, g2, s2 = cross_entropy_loss(p["ce_logits"], soft)
    rl, rg, rp = R.cross_entropy(p["ce_logits"], soft)
    assert abs(float(l2) - rl) < 1e-5 * abs(rl) and rel_err(g2, rg) < TOL


@pytest.mark.parametrize("rows,cols", [(1, 2), (100, 10), (64, 2350), (2, 26)])
def test_cross_entropy_shapes_argmax_bit_exact(nt, rows, cols):
    from net.loss import cross_entropy_loss
    rs = np.random.RandomState(rows * cols)
    logits = rs.randn(rows, cols) * 3
    logits[0, :] = 0.25                                        # exact ties: first index must win
    lab = rs.randint(0, cols, rows)
    oh = np.zeros((rows, cols))
    oh[np.arange(rows), lab] = 1
    loss, grad, sp = cross_entropy_loss(logits, oh)
    rl, rg, rp = R.cross_entropy(logits.astype(np.float32).astype(np.float64), oh)
    assert abs(float(loss) - rl) < 1e-5 * abs(rl)
    assert rel_err(grad, rg) < TOL
    host_p = sp.numpy()
    assert np.array_equal(np.argmax(sp, axis=-1), np.argmax(host_p, axis=-1))
    assert np.argmax(sp, axis=-1)[0] == 0


def test_embedding_patch_pos_golden(nt, prims):
    from net.Embedding import Embedding_layer
    from PatchEmbed import PatchEmbed_flatten
    from Position_add import Position_learnable
    p = prims
    em = Embedding_layer(9, 6, params=p["emb_E"])
    assert rel_err(em.forward(p["emb_idx"]), p["emb_y"]) < 1e-7
    assert rel_err(em.backward(p["emb_g"]), p["emb_dE"]) < TOL
    assert em.save_model()[0].dtype == np.float32               # Embedding.py:93
    flat = p["pe_params"]
    pe = PatchEmbed_flatten(8, p["pe_img"].shape, 3, True)
    pe.restore_model([[flat[:8], flat[8:16]], [flat[16:16 + 256].reshape(32, 8), flat[272:]]])
    assert rel_err(pe.forward(p["pe_img"]), p["pe_y"]) < TOL
    assert rel_err(pe.inputs, p["pe_patches"]) < 1e-7
    pl = Position_learnable(3, 8, fixed=1)
    assert rel_err(pl.posk, p["posk"]) < 1e-7
    assert pl.save_model() == []
    x = np.random.RandomState(0).randn(2, 9, 8)
    assert rel_err(pl.forward(x), x + p["posk"]) < 1e-6


def _restore_block(blk, tree):
    blk.restore_model([[a for a in pair] for pair in tree])


def test_vit_block_golden(nt, prims):
    from attention import attention_layer
    p = prims
    at = attention_layer(8, 2)
    at.restore_model(golden_tree(p, "vb_p", [[0, 0]] * 5))
    y = at.forward(p["vb_x"])
    assert rel_err(y, p["vb_y"]) < TOL
    assert rel_err(at.qkv, p["vb_qkv"]) < TOL
    assert rel_err(np.array(at.atg__), p["vb_P"]) < TOL
    dx = at.backward(p["vb_g"])
    assert rel_err(dx, p["vb_dx"]) < TOL
    assert rel_err(at.delta_qkv, p["vb_dqkv"]) < TOL
    gold = M.tree_leaves(golden_tree(p, "vb_grad", [[0, 0]] * 5))
    mine = [at.norm.gamma_delta, at.norm.beta_delta, at.qkvfc.params_delta, at.qkvfc.bias_delta,
            at.norm1.gamma_delta, at.norm1.beta_delta, at.fc0.params_delta, at.fc0.bias_delta,
            at.fc1.params_delta, at.fc1.bias_delta]
    for a, b in zip(mine, gold):
        assert rel_err(np.asarray(a).reshape(-1), b.reshape(-1)) < TOL


def test_gpt_block_golden_neg_inf_mask(nt, prims):
    from gpt.attdecoderblock import attdecoderblock_layer
    p = prims
    blk = attdecoderblock_layer(8, 2)
    blk.restore_model(golden_tree(p, "gb_p", [[0, 0]] * 6))
    y = blk.forward(p["gb_x"], M.causal_mask(6))
    assert np.all(np.isfinite(np.asarray(y)))
    assert rel_err(y, p["gb_y"]) < TOL
    assert rel_err(np.array(blk.atg__), p["gb_P"]) < TOL
    dx = blk.backward(p["gb_g"])
    assert rel_err(dx, p["gb_dx"]) < TOL
    gold = M.tree_leaves(golden_tree(p, "gb_grad", [[0, 0]] * 6))
    mine = []
    for c in blk._children():
        mine += [c.gamma_delta, c.beta_delta] if hasattr(c, "gamma") else [c.params_delta, c.bias_delta]
    for a, b in zip(mine, gold):
        assert rel_err(np.asarray(a).reshape(-1), b.reshape(-1)) < TOL
    # the -1e6 dense mask variant (attdecoderblock.py:141-148) and a dense mask on the ViT block
    y2 = blk.forward(p["gb_x"], M.causal_mask(6, neg=-1e6))
    assert rel_err(y2, p["gb_y"]) < TOL


@pytest.mark.parametrize("B,N,H,dh,causal", [(2, 49, 4, 24, False), (3, 49, 6, 64, False), (2, 5, 3, 64, True),
                                             (2, 256, 2, 64, True), (1, 100, 2, 32, False), (1, 64, 1, 64, True),
                                             (1, 128, 2, 64, False), (2, 192, 3, 64, True), (1, 68, 1, 64, True),
                                             (1, 50, 2, 24, True), (1, 70, 1, 64, False), (1, 33, 2, 16, False)])
def test_attention_core_vs_oracle(nt, B, N, H, dh, causal):
    from numpy_transformer_b200 import ops, DeviceArray
    rs = np.random.RandomState(N + dh)
    D = H * dh
    qkv = rs.randn(B, N, 3 * D) * 0.7
    dout = rs.randn(B, N, D)
    mask = M.causal_mask(N) if causal else None
    O, P, S = R.attention_fwd(qkv, H, mask)
    dq = DeviceArray.from_host(qkv)
    out, dP, dS = ops.attention_fwd(dq, H, None, ops.MASK_CAUSAL if causal else ops.MASK_NONE, want_scores=True)
    assert rel_err(out, O) < TOL and rel_err(dP, P) < TOL and rel_err(dS, S) < TOL
    dqkv = ops.attention_bwd(DeviceArray.from_host(dout), dq, dP, H, ops.MASK_CAUSAL if causal else ops.MASK_NONE)
    ref_dqkv = R.attention_bwd(dout, qkv, P, H)
    assert rel_err(dqkv, ref_dqkv) < TOL
    # with the forward's output at hand the softmax-backward row sums are dout . out (BF16x3 long-sequence kernels)
    dqkv2 = ops.attention_bwd(DeviceArray.from_host(dout), dq, dP, H, ops.MASK_CAUSAL if causal else ops.MASK_NONE, att_out=out)
    assert rel_err(dqkv2, ref_dqkv) < TOL
    if causal:
        out2, P2, _ = ops.attention_fwd(dq, H, DeviceArray.from_host(M.causal_mask(N, -1e6)), ops.MASK_DENSE)
        assert rel_err(out2, O) < TOL


@pytest.mark.parametrize("N,causal", [(49, False), (5, True), (64, True)])
def test_attention_bf16_kernels_single_block(nt, N, causal, monkeypatch):
    """Single-block BF16x3 attention kernels (csrc/attention_bf16.cu, N <= 64), forced on also below the default
    threshold of 16 tokens (where the TF32x3 kernels of test_attention_core_vs_oracle run by default)."""
    from numpy_transformer_b200 import ops, DeviceArray
    monkeypatch.setenv("NTB200_ATT_BF16_MIN_N", "1")
    rs = np.random.RandomState(N)
    B, H, dh = 2, 3, 64
    qkv, dout = rs.randn(B, N, 3 * H * dh) * 0.7, rs.randn(B, N, H * dh)
    O, P, S = R.attention_fwd(qkv, H, M.causal_mask(N) if causal else None)
    dq = DeviceArray.from_host(qkv)
    out, dP, dS = ops.attention_fwd(dq, H, None, ops.MASK_CAUSAL if causal else ops.MASK_NONE, want_scores=True)
    assert rel_err(out, O) < TOL and rel_err(dP, P) < TOL and rel_err(dS, S) < TOL
    dqkv = ops.attention_bwd(DeviceArray.from_host(dout), dq, dP, H, ops.MASK_CAUSAL if causal else ops.MASK_NONE)
    assert rel_err(dqkv, R.attention_bwd(dout, qkv, P, H)) < TOL


def test_empty_batch_and_errors(nt):
    from net.fullconnect import fclayer
    from net.activation import Softmax
    fc = fclayer(4, 3, True)
    y = fc.forward(np.zeros((0, 4)))
    assert y.shape == (0, 3)
    with pytest.raises(AssertionError):
        fc.forward(np.zeros((2, 5)))
    sm = Softmax()
    sm.forward(np.zeros((2, 3, 4)), axis=-1)
    with pytest.raises(RuntimeError):
        sm.backward(np.zeros((2, 3, 4)), np.zeros((2, 3, 4)))
    with pytest.raises(nt.NtError):
        nt.lib.nt_cross_entropy(None, None, None, 1.0, None, None, None, None, None, 0, 0)


@pytest.mark.parametrize("M_,K,N", [(1, 384, 384), (2, 1536, 384), (4, 384, 1152), (5, 100, 10), (8, 384, 2350), (3, 67, 1537),
                                    (10, 192, 576), (16, 768, 192), (13, 960, 26)])
def test_skinny_linear_forward_and_layernorm_fusion(nt, M_, K, N):
    """<= 16-row Linear forward (linear_skinny_kernel: every column-width variant, vector and scalar weight loads, bias /
    activation / pre-activation / residual epilogue) and its LayerNorm-fused decode form nt_linear_ln_skinny, vs the oracle."""
    from numpy_transformer_b200 import ops, DeviceArray
    from numpy_transformer_b200._lib import lib
    rs = np.random.RandomState(M_ + K + N)
    x, W, b, res = rs.randn(M_, K), rs.randn(K, N) / np.sqrt(K), rs.randn(N), rs.randn(M_, N)
    dx, dW, db, dres = (DeviceArray.from_host(a) for a in (x, W, b, res))
    before = nt.launch_count()
    y, pre = ops.linear_fwd(dx, dW, db, ops.ACT_GELU, want_pre=True, residual=dres)
    assert nt.launch_count() - before == 1
    ref_pre = R.linear_fwd(x, W, b)
    assert rel_err(pre, ref_pre) < 1e-5 and rel_err(y, R.gelu_fwd(ref_pre) + res) < 1e-5
    assert rel_err(ops.linear_fwd(dx, dW, None, ops.ACT_RELU), np.maximum(R.linear_fwd(x, W), 0)) < 1e-5
    g, be = rs.randn(K) * 0.5 + 1, rs.randn(K) * 0.1
    out, dg, dbe = DeviceArray((M_, N)), DeviceArray.from_host(g), DeviceArray.from_host(be)
    lib.nt_linear_ln_skinny(dx.ptr, dg.ptr, dbe.ptr, dW.ptr, db.ptr, out.ptr, M_, K, N, ops.ACT_RELU, dres.ptr)
    u = R.layernorm_fwd(x, g, be)[0]
    assert rel_err(out, np.maximum(R.linear_fwd(u, W, b), 0) + res) < 1e-5


@pytest.mark.parametrize("M_,K,N", [(1, 192, 576), (5, 384, 26), (10, 768, 192), (16, 100, 30), (10, 192, 768), (3, 2352, 384)])
def test_skinny_linear_backward(nt, M_, K, N):
    """<= 16-row Linear backward (linear_skinny_bwd_input_kernel / linear_skinny_bwd_params_kernel: the 10-row layers of the
    gpt_character configuration): dX with the fused act'(pre) and addend, dW / db accumulating over two calls
    (fullconnect.py:118-127), vector and scalar paths, one launch each, vs the oracle."""
    from numpy_transformer_b200 import ops, DeviceArray
    rs = np.random.RandomState(M_ * 7 + K + N)
    x, W, dy = rs.randn(M_, K), rs.randn(K, N) / np.sqrt(K), rs.randn(M_, N)
    pre, add = rs.randn(M_, K), rs.randn(M_, K)
    dW_, ddy, dpre, dadd, dx_ = (DeviceArray.from_host(a) for a in (W, dy, pre, add, x))
    before = nt.launch_count()
    g = ops.linear_bwd_input(ddy, dW_, ops.ACT_RELU, dpre, dadd)
    assert nt.launch_count() - before == 1
    ref = dy @ W.T
    assert rel_err(g, ref * (pre > 0) + add) < 1e-5
    assert rel_err(ops.linear_bwd_input(ddy, dW_), ref) < 1e-5
    assert rel_err(ops.linear_bwd_input(ddy, dW_, ops.ACT_GELU, dpre), R.gelu_bwd(ref, pre)) < 1e-5
    gw, gb = DeviceArray.zeros((K, N)), DeviceArray.zeros((N,))
    before = nt.launch_count()
    ops.linear_bwd_params(dx_, ddy, gw, gb)
    assert nt.launch_count() - before == 1
    ops.linear_bwd_params(dx_, ddy, gw, gb)
    assert rel_err(gw, 2 * (x.T @ dy)) < 1e-5 and rel_err(gb, 2 * dy.sum(0)) < 1e-5
