"""GPU parity of the fused ViT-block kernels (csrc/vit_block.cu, nt_vit_blocks_fwd/bwd) against the oracle's
restatement of attention.py:20-89.  Tolerance: max|a-b|/max|b| <= 1e-4 per tensor (fp32-class path)."""
import os
import numpy as np
import pytest
from conftest import rel_err
from oracle import models as M

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    nt.set_gemm_mode(1)
    return nt


def _block_tree(layer):
    return [[np.asarray(a, dtype=np.float64).reshape(-1) if np.asarray(a).ndim != 2 else np.asarray(a, dtype=np.float64)
             for a in pair] for pair in layer.save_model()]


def _grads(layer):
    return [layer.norm.gamma_delta, layer.norm.beta_delta, layer.qkvfc.params_delta, layer.qkvfc.bias_delta,
            layer.norm1.gamma_delta, layer.norm1.beta_delta, layer.fc0.params_delta, layer.fc0.bias_delta,
            layer.fc1.params_delta, layer.fc1.bias_delta]


def _randomise(layer, rng):
    """LayerNorm affine parameters away from (1, 0) so their gradients are exercised."""
    for ln in (layer.norm, layer.norm1):
        ln.gamma.set(1.0 + 0.3 * rng.standard_normal(ln.gamma.size))
        ln.beta.set(0.2 * rng.standard_normal(ln.beta.size))


@pytest.mark.parametrize("D,H,B,N", [(32, 2, 3, 49), (64, 4, 2, 49), (96, 4, 5, 49), (96, 6, 2, 33), (64, 2, 3, 16),
                                     (96, 4, 2, 64), (32, 2, 1, 1), (96, 4, 3, 7)])
def test_fused_block_layer_api_vs_oracle(nt, D, H, B, N):
    from attention import attention_layer
    from numpy_transformer_b200 import fused_vit
    assert fused_vit.supported(N, D, H)
    rng = np.random.default_rng(D + H + N)
    np.random.seed(7)
    at = attention_layer(D, H)
    _randomise(at, rng)
    blk = _block_tree(at)
    x = rng.standard_normal((B, N, D))
    gin = rng.standard_normal((B, N, D))
    ref_y, cache = M.vit_block_fwd(x, blk, H)
    ref_dx, ref_g, extra = M.vit_block_bwd(gin, blk, cache, H)
    before = nt.launch_count()
    y = at.forward(x)
    assert at._fused and nt.launch_count() - before == 2          # weight split + the fused forward
    assert rel_err(y, ref_y) < TOL
    assert rel_err(at.qkv, cache["qkv"]) < TOL
    assert rel_err(np.array(at.atg__), cache["P"]) < TOL
    assert rel_err(at.input1, cache["h"]) < TOL
    before = nt.launch_count()
    dx = at.backward(gin)
    assert nt.launch_count() - before == 1
    assert rel_err(dx, ref_dx) < TOL
    for i, (a, b_) in enumerate(zip(_grads(at), M.tree_leaves(ref_g))):
        assert rel_err(np.asarray(a).reshape(-1), b_.reshape(-1)) < TOL, "grad leaf %d" % i
    # accumulate-until-setzero (fullconnect.py:120-125): a second backward doubles every gradient
    at.backward(gin)
    for i, (a, b_) in enumerate(zip(_grads(at), M.tree_leaves(ref_g))):
        assert rel_err(np.asarray(a).reshape(-1), 2 * b_.reshape(-1)) < TOL, "accumulated grad leaf %d" % i
    at.setzero()
    assert not any(np.any(np.asarray(a)) for a in _grads(at))


@pytest.mark.parametrize("D,H,L", [(96, 4, 6), (32, 2, 3)])
def test_fused_stack_one_launch_vs_oracle(nt, D, H, L):
    from attention import attention_layer
    from numpy_transformer_b200 import fused_vit
    from numpy_transformer_b200.device import as_device
    rng = np.random.default_rng(11)
    np.random.seed(3)
    layers = [attention_layer(D, H) for _ in range(L)]
    for l in layers:
        _randomise(l, rng)
    B, N = 4, 49
    x = rng.standard_normal((B, N, D))
    gin = rng.standard_normal((B, N, D)) / N
    xs, caches = [x], []
    for l in layers:
        y, c = M.vit_block_fwd(xs[-1], _block_tree(l), H)
        xs.append(y)
        caches.append(c)
    d = gin
    ref_grads = [None] * L
    for i in reversed(range(L)):
        d, ref_grads[i], _ = M.vit_block_bwd(d, _block_tree(layers[i]), caches[i], H)
    before = nt.launch_count()
    y = fused_vit.forward_stack(layers, as_device(x))
    dx = fused_vit.backward_stack(layers, as_device(gin))
    assert nt.launch_count() - before == 3                          # split, forward, backward for the whole stack
    assert rel_err(y, xs[-1]) < TOL
    assert rel_err(dx, d) < TOL
    for i, l in enumerate(layers):
        assert rel_err(l._out, xs[i + 1]) < TOL
        assert np.allclose(np.asarray(l.P).sum(-1), 1.0, atol=1e-5)
        for j, (a, b_) in enumerate(zip(_grads(l), M.tree_leaves(ref_grads[i]))):
            assert rel_err(np.asarray(a).reshape(-1), b_.reshape(-1)) < TOL, "block %d grad leaf %d" % (i, j)


def test_fused_and_unfused_paths_agree(nt):
    from attention import attention_layer
    rng = np.random.default_rng(5)
    x = rng.standard_normal((3, 49, 96))
    gin = rng.standard_normal((3, 49, 96))
    outs = []
    for fused in ("1", "0"):
        os.environ["NTB200_FUSED_VIT"] = fused
        try:
            np.random.seed(1)
            at = attention_layer(96, 4)
            y = np.asarray(at.forward(x))
            assert bool(at._fused) == (fused == "1")
            dx = np.asarray(at.backward(gin))
            outs.append([y, dx] + [np.asarray(a) for a in _grads(at)])
        finally:
            os.environ.pop("NTB200_FUSED_VIT", None)
    for a, b_ in zip(*outs):
        assert rel_err(a, b_) < TOL


@pytest.mark.parametrize("use_graph", [False, True])
def test_trainstep_uses_fused_stack(nt, use_graph):
    """TrainStep groups the run of attention_layer objects into one forward and one backward launch and still
    reproduces the oracle's full training step (transformer_of_image.py:141-159)."""
    from numpy_transformer_b200 import trainer
    cfg = M.ViTConfig(batch=6, embed_dim=32, heads=2, blocks=3)
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    layers = trainer.build_vit_layers(6, embed_dim=32, heads=2, blocks=3, seed=0)
    ts = trainer.TrainStep(layers, "vit", images.shape, lr=0.02, use_graph=use_graph)
    for step in range(4):
        ref = M.vit_step(params, images, onehot, 0.02, cfg)
        loss, amax = ts.step(images, labels, want_argmax=True)
        assert abs(loss - ref["loss"]) < TOL * abs(ref["loss"]), step
        assert np.array_equal(amax, ref["argmax"])
        params = ref["params"]
        mine = M.tree_leaves(trainer.save_walk(layers))
        for i, (a, b_) in enumerate(zip(mine, M.tree_leaves(params))):
            assert rel_err(a.reshape(-1), b_.reshape(-1)) < TOL, "step %d leaf %d" % (step, i)
    assert all(getattr(l, "_fused", False) for l in layers[2:5])
    if use_graph:
        assert ts.graph is not None and ts.graph_kernels < 40
