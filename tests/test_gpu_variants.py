"""GPU parity for the non-README switches of the mirror and the DeviceArray
NumPy interop the reference's trainers rely on.  Checked against direct NumPy restatements."""
import numpy as np
import pytest
from conftest import rel_err
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


def test_gpt_linear_layer_head(nt):
    """gpt/gpt_linear.py:6-44  FC0(D->3D) -> LN -> FC1."""
    from gpt.gpt_linear import gpt_linear_layer
    np.random.seed(3)
    head = gpt_linear_layer(16, 11)
    (W0, b0), (W1, b1), (g, be) = head.save_model()
    x = np.random.randn(5, 7, 16)
    y = head.forward(x)
    h0 = R.linear_fwd(x, W0, b0)
    h1, xhat, var = R.layernorm_fwd(h0, g.reshape(-1), be.reshape(-1))
    ref = R.linear_fwd(h1, W1, b1)
    assert rel_err(y, ref) < TOL
    dy = np.random.randn(5, 7, 11)
    dx = head.backward(dy)
    d1, dW1, db1 = R.linear_bwd(dy, h1, W1)
    d0, dg, dbeta = R.layernorm_bwd(d1, xhat, var, g.reshape(-1))
    rdx, dW0, db0 = R.linear_bwd(d0, x, W0)
    assert rel_err(dx, rdx) < TOL
    assert rel_err(head.fc0.params_delta, dW0) < TOL and rel_err(head.fc1.params_delta, dW1) < TOL
    assert rel_err(head.norm.gamma_delta, dg) < TOL and rel_err(head.norm.beta_delta, dbeta) < TOL


def test_position_fixed_and_learnable(nt):
    """gpt/FixedEmbed.py:11-40 lookup; Position_add.py learnable branch (:20-39)."""
    from gpt.FixedEmbed import Position_Fixed
    from Position_add import Position_learnable
    pf = Position_Fixed(10, 8)
    idx = np.array([[0, 3, 9], [1, 1, 2]])
    assert np.array_equal(pf.forward(idx), R.posk_table(10, 8)[0][idx])
    assert pf.save_model() == []
    np.random.seed(1)
    pl = Position_learnable(3, 8, fixed=0)
    param0 = np.asarray(pl.param).copy()
    x = np.random.randn(4, 9, 8)
    assert rel_err(pl.forward(x), x + param0) < 1e-6
    d = np.random.randn(4, 9, 8)
    out = pl.backward(d)
    assert rel_err(out, d) < 1e-7
    assert rel_err(pl.param_delta, d.sum(0)[None]) < 1e-5
    pl.update(0.1)
    assert rel_err(pl.param, param0 - 0.1 * d.sum(0)[None]) < 1e-5
    pl.setzero()
    assert not np.any(np.asarray(pl.param_delta))
    assert pl.save_model()[0].shape == (1, 9, 8)


def test_classify_cls_token_branch(nt):
    """classify.py:12-13,35-38: cls_token head consumes token 0 and scatters its gradient back."""
    from classify import classify_layer
    np.random.seed(5)
    cl = classify_layer(8, 3, 2, 5, cls_token=True)
    (W0, b0), (W1, b1) = cl.save_model()
    x = np.random.randn(3, 4, 8)
    y = cl.forward(x[:, 0])
    h = R.relu_fwd(R.linear_fwd(x[:, 0], W0, b0))
    assert rel_err(y, R.linear_fwd(h, W1, b1)) < TOL
    dy = np.random.randn(3, 5)
    dx = np.asarray(cl.backward(dy))
    assert dx.shape == (3, 4, 8) and not np.any(dx[:, 1:])
    d1, _, _ = R.linear_bwd(dy, h, W1)
    d0, _, _ = R.linear_bwd(R.relu_bwd(d1, R.linear_fwd(x[:, 0], W0, b0)), x[:, 0], W0)
    assert rel_err(dx[:, 0], d0) < TOL


def test_embedding_layer_api(nt):
    """net/Embedding.py:57-90: explicit `flatten` override, normdelta, SGD update, duplicates accumulate."""
    from net.Embedding import Embedding_layer
    np.random.seed(2)
    em = Embedding_layer(7, 4)
    E = np.asarray(em.params).copy()
    idx = np.array([[1, 1, 6], [0, 1, 3]])
    em.forward(idx)
    g = np.random.randn(2, 3, 4)
    em.backward(g)
    ref = R.embedding_bwd(idx, g, 7)
    assert rel_err(em.getdelta(), ref) < 1e-6
    em.normdelta(2.0, 4.0)
    assert rel_err(em.delta, ref * 0.5) < 1e-6
    em.update(0.1)
    assert rel_err(em.params, E - 0.1 * ref * 0.5) < 1e-6
    em.setzero()
    em.backward(g, flatten=np.array([2, 2, 2, 2, 2, 2]))
    assert rel_err(np.asarray(em.delta)[2], g.reshape(-1, 4).sum(0)) < 1e-6


def test_device_array_numpy_interop(nt):
    """What the trainers do with layer outputs: reshape, len, indexing, argmax, zeros_like, format, arithmetic."""
    from numpy_transformer_b200 import DeviceArray
    a = np.arange(24, dtype=np.float64).reshape(2, 3, 4)
    d = DeviceArray.from_host(a)
    assert d.shape == (2, 3, 4) and len(d) == 2 and d.ndim == 3 and d.size == 24
    assert np.array_equal(np.asarray(d), a) and np.asarray(d).dtype == np.float64
    r = np.reshape(d, (-1, 4))
    assert isinstance(r, DeviceArray) and r.shape == (6, 4)
    assert np.array_equal(d[:, 0], a[:, 0])
    assert np.zeros_like(d).shape == (2, 3, 4)
    assert np.array_equal(np.argmax(r, axis=-1), np.argmax(a.reshape(6, 4), -1))
    s = DeviceArray.from_host(np.array(2.5)).reshape(())
    assert "{:.6f}".format(s) == "2.500000" and float(s) == 2.5 and s + 1 == 3.5 and 2 * s == 5.0
    total = 0
    total += s
    assert total == 2.5
    e = d + d
    assert isinstance(e, DeviceArray) and np.array_equal(np.asarray(e), 2 * a)
    assert np.array_equal(np.asarray(d.astype(np.float32)), a.astype(np.float32))
    import copy
    c = copy.deepcopy(d)
    c.fill_zero()
    assert np.array_equal(np.asarray(d), a)


def test_return_attention_mode(nt):
    """gpt/attdecoderblock.py:58-59,73-74: return_attention hands back (atg__, att_col) as [batch][head] lists."""
    from gpt.attdecoderblock import attdecoderblock_layer
    from oracle import models as M
    np.random.seed(9)
    blk = attdecoderblock_layer(16, 2, return_attention=True)
    x = np.random.randn(2, 6, 16)
    atg, att = blk.forward(x, M.causal_mask(6))
    assert len(atg) == 2 and len(atg[0]) == 2 and atg[0][0].shape == (6, 6)
    tree = blk.save_model()
    sq = [[np.asarray(a).reshape(-1) if np.asarray(a).ndim == 3 else np.asarray(a) for a in pair] for pair in tree]
    _, c = M.gpt_block_fwd(x, sq, 2, M.causal_mask(6))
    assert rel_err(np.array(atg), c["P"]) < TOL
    from oracle import numpy_ref as RR
    _, _, S = RR.attention_fwd(c["qkv"], 2, None)
    assert rel_err(np.array(att), S) < TOL


def test_progressive_block_growth_restore(nt):
    """gpt/gpt_train_potry3000.py:212-231: a checkpoint of an N-block model restored into an (N+1)-block model.  The
    reference's loop relies on restore_model RAISING when the new block is handed the final LayerNorm's [gamma, beta]
    (its restore indexes models[0..5]), then seeds the block from the previous entry.  Run that loop verbatim on the
    drop-in objects, and the same through trainer.restore_walk(grow=True)."""
    import pickle
    from numpy_transformer_b200 import trainer
    small = trainer.build_gpt_layers(2, 8, 32, 2, 2, 11, adam=True, seed=0)
    x = np.random.RandomState(0).randint(0, 11, (2, 8))
    for l in small:                                     # one forward: LayerNorm checkpoints take their rank-expanded shape
        x = l.forward(x, "causal") if type(l).__name__ == "attdecoderblock_layer" else l.forward(x)
    tree = trainer.save_walk(small)
    tree.append(1234)                                   # the trainer's step counter (:370)
    models = pickle.loads(pickle.dumps(tree))

    def grown_layers():
        return trainer.build_gpt_layers(2, 8, 32, 2, 3, 11, adam=True, seed=7)

    # (1) the reference's own loop, verbatim
    layers = grown_layers()
    jk, cnt, num_attention = models[-1], 0, 0
    for l in layers:
        k = dir(l)
        nam = l.__module__
        if 'attdecoderblock' in nam:
            num_attention += 1
        if 'restore_model' in k and 'save_model' in k:
            try:
                l.restore_model(models[cnt])
            except:                                     # noqa: E722  (as in the reference)
                jk = 0
                l.restore_model(models[cnt - 1])
                continue
            cnt += 1
    assert num_attention == 3 and jk == 0 and cnt == 5
    got = trainer.save_walk(layers)
    want = [tree[0], tree[1], tree[2], tree[2], tree[3], tree[4]]      # the new block is a copy of the block before it
    for a, b in zip(M.tree_leaves(got), M.tree_leaves(want)):
        assert np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1))
    # (2) a failed restore leaves the layer untouched (validated before anything is overwritten)
    probe = grown_layers()
    before = [np.asarray(a).copy() for a in M.tree_leaves(probe[3].save_model())]
    with pytest.raises((IndexError, ValueError)):
        probe[3].restore_model(models[3])               # the final LayerNorm's [gamma, beta]
    for a, b in zip(M.tree_leaves(probe[3].save_model()), before):
        assert np.array_equal(np.asarray(a), b)
    with pytest.raises((IndexError, ValueError)):
        probe[3].restore_model(models[2][:4])           # a block tree cut short
    # (3) the library's walk does the same and reports it
    layers2 = grown_layers()
    step, grown = trainer.restore_walk(layers2, models, grow=True)
    assert grown == 1 and step is None
    for a, b in zip(M.tree_leaves(trainer.save_walk(layers2)), M.tree_leaves(want)):
        assert np.array_equal(np.asarray(a).reshape(-1), np.asarray(b).reshape(-1))
    step, grown = trainer.restore_walk(trainer.build_gpt_layers(2, 8, 32, 2, 2, 11, adam=True, seed=9), models, grow=True)
    assert (step, grown) == (1234, 0)
    # the grown model trains
    ts = trainer.TrainStep(layers2, "gpt", (2, 8), lr=1e-3, adam=True)
    assert np.isfinite(ts.step(np.random.RandomState(1).randint(0, 11, (2, 8)), np.random.RandomState(2).randint(0, 11, (2, 8))))
    ts.close()
