"""BF16x3 tcgen05 engine for the large Linear shapes (csrc/split_bf16.cu + gemm_tc.cu bf16 mode) against fp64 NumPy.
Tolerance: the fp32-class bar, max|a-b|/max|b| <= 1e-4 (measured 3e-6 .. 1.2e-5).  The split itself must reconstruct
x to 2^-16 relative and reproduce its layout contract bit for bit."""
import os
import numpy as np
import pytest
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    nt.set_gemm_mode(1)
    old = os.environ.get("NTB200_BF16X3_MIN_ROWS")
    os.environ["NTB200_BF16X3_MIN_ROWS"] = "1"          # force the engine for the small test shapes
    yield nt
    if old is None:
        os.environ.pop("NTB200_BF16X3_MIN_ROWS", None)
    else:
        os.environ["NTB200_BF16X3_MIN_ROWS"] = old


def _bf16_planes(arr, rows, ld):
    """device bf16 plane (stored two per fp32 word) -> float64 [rows, ld]"""
    u16 = arr.numpy().view(np.uint16).reshape(rows, ld)
    return (u16.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


@pytest.mark.parametrize("rows,cols", [(64, 64), (130, 72), (49, 16), (1000, 384), (7, 8)])
def test_split_layouts(nt, rows, cols):
    from numpy_transformer_b200 import ops, DeviceArray
    rs = np.random.RandomState(rows + cols)
    x = (rs.randn(rows, cols) * 10 ** rs.uniform(-3, 3, (rows, cols))).astype(np.float32)
    dx = DeviceArray.from_host(x)
    hi, lo, ld = ops.split_bf16(dx, rows, cols)
    h, l = _bf16_planes(hi, rows, ld), _bf16_planes(lo, rows, ld)
    assert np.max(np.abs(h + l - x) / np.abs(x)) < 2.0 ** -15
    hit, lot, ldt = ops.split_bf16(dx, rows, cols, transpose=True)
    assert ldt % 8 == 0 and ldt >= rows
    ht, lt = _bf16_planes(hit, cols, ldt), _bf16_planes(lot, cols, ldt)
    assert np.array_equal(ht[:, :rows], h.T) and np.array_equal(lt[:, :rows], l.T)
    if cols % 4 == 0:
        (h2, l2), (h2t, l2t, ld2) = ops.split_bf16_both(dx, rows, cols)
        assert np.array_equal(_bf16_planes(h2, rows, cols), h) and np.array_equal(_bf16_planes(l2, rows, cols), l)
        assert np.array_equal(_bf16_planes(h2t, cols, ld2)[:, :rows], h.T)
        assert np.array_equal(_bf16_planes(l2t, cols, ld2)[:, :rows], l.T)


@pytest.mark.parametrize("M,K,N", [(128, 64, 128), (300, 384, 1152), (4900, 96, 288), (777, 200, 264), (2049, 768, 384),
                                   (100, 4704, 96), (513, 104, 40)])
def test_linear_bf16x3_three_gemms(nt, M, K, N):
    from numpy_transformer_b200 import ops, DeviceArray
    from oracle import numpy_ref as R
    assert ops.bf16x3_ok(M, K, N)
    rs = np.random.RandomState(M * 31 + K * 7 + N)
    x, w, b, dy = rs.randn(M, K), rs.randn(K, N) / np.sqrt(K), rs.randn(N), rs.randn(M, N)
    res = rs.randn(M, N)
    dx_, dw_, db_, ddy, dres = (DeviceArray.from_host(a) for a in (x, w, b, dy, res))
    before = nt.launch_count()
    keep = []
    y, pre = ops.linear_fwd(dx_, dw_, db_, ops.ACT_GELU, want_pre=True, residual=dres, keep_xt=keep)
    assert nt.launch_count() - before == 3 and len(keep) == 1          # split(x, both layouts), split(w^T), GEMM
    ref_pre = x @ w + b
    assert rel_err(pre, ref_pre) < TOL
    assert rel_err(y, R.gelu_fwd(ref_pre) + res) < TOL
    assert rel_err(ops.linear_bwd_input(ddy, dw_), dy @ w.T) < TOL
    pre_k, add_k = DeviceArray.from_host(rs.randn(M, K)), DeviceArray.from_host(rs.randn(M, K))
    dx2 = ops.linear_bwd_input(ddy, dw_, ops.ACT_RELU, pre_k, add_k)
    assert rel_err(dx2, (dy @ w.T) * (np.asarray(pre_k) > 0) + np.asarray(add_k)) < TOL
    g, gb = DeviceArray.zeros((K, N)), DeviceArray.zeros((N,))
    ops.linear_bwd_params(dx_, ddy, g, gb, xt_split=keep[0])            # transposed split of x cached by the forward
    ops.linear_bwd_params(dx_, ddy, g, gb)                              # accumulates (+=), splits x itself
    assert rel_err(g, 2 * (x.T @ dy)) < TOL
    assert rel_err(gb, 2 * dy.sum(0)) < TOL


def test_fclayer_uses_engine_and_matches_tf32x3_path(nt):
    """The drop-in layer picks the engine by shape; both engines of mode 1 agree to the fp32-class bar."""
    from net.fullconnect import fclayer
    rs = np.random.RandomState(3)
    x, dy = rs.randn(2304, 96), rs.randn(2304, 192)
    outs = []
    for min_rows in ("1", "1000000"):
        os.environ["NTB200_BF16X3_MIN_ROWS"] = min_rows
        np.random.seed(0)
        fc = fclayer(96, 192, True)
        y = np.asarray(fc.forward(x))
        dx = np.asarray(fc.backward(dy))
        outs.append((y, dx, np.asarray(fc.params_delta), np.asarray(fc.bias_delta)))
    os.environ["NTB200_BF16X3_MIN_ROWS"] = "1"
    for a, b_ in zip(*outs):
        assert rel_err(a, b_) < TOL


def test_vit_per_op_path_with_producer_splits(nt):
    """Whole ViT step on the per-op path (D = 128 is outside the fused block kernels) with the BF16x3 engine forced:
    LayerNorm, GEMM-epilogue and attention kernels hand their bf16 planes to the next Linear (DeviceArray._split)."""
    from numpy_transformer_b200 import trainer
    from oracle import models as M
    cfg = M.ViTConfig(batch=4, embed_dim=128, heads=2, blocks=2)
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    layers = trainer.build_vit_layers(4, embed_dim=128, heads=2, blocks=2, seed=0)
    ts = trainer.TrainStep(layers, "vit", images.shape, lr=0.02)
    for step in range(3):
        ref = M.vit_step(params, images, onehot, 0.02, cfg)
        loss, amax = ts.step(images, labels, want_argmax=True)
        assert abs(loss - ref["loss"]) < TOL * abs(ref["loss"]), step
        assert np.array_equal(amax, ref["argmax"])
        params = ref["params"]
        for i, (a, b_) in enumerate(zip(M.tree_leaves(trainer.save_walk(layers)), M.tree_leaves(params))):
            assert rel_err(a.reshape(-1), b_.reshape(-1)) < TOL, "step %d leaf %d" % (step, i)
    blk = layers[2]
    assert not getattr(blk, "_fused", False)
    assert blk.out._split is not None and blk.out1._split is not None and blk.out2._split is not None   # LN, LN, GELU epilogue
    assert blk.delta_qkv._split is not None                                                            # attention backward


def test_gpt_step_with_producer_splits(nt):
    """GPT blocks (dh = 64, T = 128) with the BF16x3 engine forced: loss of the first step and after two SGD updates
    (which carries every gradient) against the oracle; attention forward / backward and LayerNorm backward emit the
    planes their consumers read."""
    from numpy_transformer_b200 import trainer
    from oracle import models as M
    cfg = M.GPTConfig(batch=2, context=128, embed_dim=128, heads=2, blocks=2, vocab=48)
    params = M.gpt_init(cfg, 0)
    tokens, labels = M.synthetic_gpt_batch(cfg, 1)
    layers = trainer.build_gpt_layers(2, 128, 128, 2, 2, 48, adam=False, seed=0)
    ts = trainer.TrainStep(layers, "gpt", tokens.shape, lr=0.05, adam=False)
    for step in range(3):
        ref = M.gpt_step(params, None, tokens, labels, 0.05, cfg, adam=False)
        loss = ts.step(tokens, labels)
        assert abs(loss - ref["loss"]) < TOL * abs(ref["loss"]), step
        params = ref["params"]
    blk = layers[1]
    assert blk.rkk._split is not None and blk.delta_qkv._split is not None and blk.out0._split is not None
