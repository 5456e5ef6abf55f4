"""The tcgen05 / TMEM fused ViT block stack (csrc/vit_block_tc.cu) against the oracle (oracle/models.py restates
attention.py:20-89): every tensor the forward leaves behind and every gradient of the backward, per tensor <= 1e-4."""
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


def _stack(nt, B, N, L, seed):
    from attention import attention_layer
    np.random.seed(seed)
    layers = [attention_layer(96, 4) for _ in range(L)]
    # biases / LayerNorm parameters away from their trivial initial values
    rs = np.random.RandomState(seed + 1)
    for l in layers:
        for ln in (l.norm, l.norm1):
            ln.gamma.set(1.0 + 0.2 * rs.randn(96))
            ln.beta.set(0.1 * rs.randn(96))
    x = rs.randn(B, N, 96)
    blks = [[[np.asarray(c.gamma, dtype=np.float64), np.asarray(c.beta, dtype=np.float64)] if hasattr(c, "gamma") else
             [np.asarray(c.params, dtype=np.float64), np.asarray(c.bias_params, dtype=np.float64)] for c in l._children()]
            for l in layers]
    return layers, blks, x


@pytest.mark.parametrize("B,N,L", [(3, 49, 2), (5, 64, 1), (2, 17, 3), (100, 49, 6)])
def test_tc_forward_vs_oracle(nt, monkeypatch, B, N, L):
    from numpy_transformer_b200 import fused_vit
    monkeypatch.setenv("NTB200_VIT_TC", "1")
    assert fused_vit.tc_supported(N, 96, 4)
    layers, blks, x = _stack(nt, B, N, L, seed=B + N)
    out = fused_vit.forward_stack(layers, nt.as_device(x))
    assert layers[0]._fused == "tc"
    ref = x
    for l, blk in zip(layers, blks):
        ref, c = M.vit_block_fwd(ref, blk, 4)
        assert rel_err(l.qkv, c["qkv"]) < TOL
        assert rel_err(l.P, c["P"]) < TOL
        assert rel_err(l.input1, c["h"]) < TOL
        assert rel_err(l._dgelu, c["pre"]) < TOL
        assert rel_err(l._out, ref) < TOL
    assert rel_err(out, ref) < TOL
    # the softmax is optional: same outputs without materialising it
    out2 = fused_vit.forward_stack(layers, nt.as_device(x), want_P=False)
    assert layers[0].P is None and np.array_equal(np.asarray(out2), np.asarray(out))


@pytest.mark.parametrize("B,N,L", [(3, 49, 2), (4, 64, 1), (2, 17, 3), (100, 49, 6)])
def test_tc_backward_vs_oracle(nt, monkeypatch, B, N, L):
    """dx and every parameter gradient of the stack (accumulated over two backward calls: the reference's `+=` until
    setzero) against the oracle's hand-derived adjoint (attention.py:57-89)."""
    from numpy_transformer_b200 import fused_vit
    monkeypatch.setenv("NTB200_VIT_TC", "1")
    layers, blks, x = _stack(nt, B, N, L, seed=7 * B + N)
    rs = np.random.RandomState(5)
    dout = rs.randn(B, N, 96) / np.sqrt(96)
    for want_P in (True, False):
        out = fused_vit.forward_stack(layers, nt.as_device(x), want_P=want_P)
        dx = fused_vit.backward_stack(layers, nt.as_device(dout))
    # oracle
    ref, caches = x, []
    for blk in blks:
        ref, c = M.vit_block_fwd(ref, blk, 4)
        caches.append(c)
    d, grads = dout, []
    for blk, c in zip(reversed(blks), reversed(caches)):
        d, gr, _ = M.vit_block_bwd(d, blk, c, 4)
        grads.insert(0, gr)
    assert rel_err(dx, d) < TOL
    for li, (l, gr) in enumerate(zip(layers, grads)):
        mine = [[c.gamma_delta, c.beta_delta] if hasattr(c, "gamma") else [c.params_delta, c.bias_delta] for c in l._children()]
        for i, (a_, b_) in enumerate(zip(M.tree_leaves(mine), M.tree_leaves(gr))):
            e = rel_err(np.asarray(a_).reshape(-1), 2.0 * np.asarray(b_).reshape(-1))      # two accumulated passes
            assert e < TOL, "block %d leaf %d rel err %.2e" % (li, i, e)
        for c in l._children():
            c.setzero()
