"""KV-cache generation (numpy_transformer_b200/generate.py) against the reference's way of sampling:
a forward over the whole window for every new token (gpt/gpt_predict_poetrythree.py:108-121)."""
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
    return nt


@pytest.mark.parametrize("D,H", [(32, 2), (128, 2)])
def test_kv_cache_matches_full_window(nt, D, H):
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.generate import KVGenerator
    from oracle import models as M
    B, T, V = 2, 24, 13
    layers = trainer.build_gpt_layers(B, T, D, H, 2, V, adam=False, seed=0)
    gen = KVGenerator(layers)
    rs = np.random.RandomState(0)
    toks = rs.randint(0, V, (B, 6))
    # oracle check of the full-window logits themselves (one position)
    cfg = M.GPTConfig(batch=B, context=6, embed_dim=D, heads=H, blocks=2, vocab=V)
    ref_logits, _ = M.gpt_forward(M.tree_unflatten(M.gpt_init(cfg, 0), M.tree_leaves(trainer.save_walk(layers))), toks, cfg,
                                  M.causal_mask(6))
    logits = gen.prefill(toks)
    assert rel_err(logits, np.asarray(ref_logits).reshape(B, 6, V)[:, -1]) < TOL
    params = M.tree_unflatten(M.gpt_init(cfg, 0), M.tree_leaves(trainer.save_walk(layers)))
    seq = toks.copy()
    for _ in range(T - 6):
        nxt = np.argmax(logits, -1)
        seq = np.concatenate([seq, nxt[:, None]], 1)
        logits = gen.step(nxt)
        # EVERY cached step against the oracle's forward over the whole window so far (the reference's procedure,
        # gpt_predict_poetrythree.py:108-121), not against this repo's own full-window path
        n = seq.shape[1]
        ocfg = M.GPTConfig(batch=B, context=n, embed_dim=D, heads=H, blocks=2, vocab=V)
        want = np.asarray(M.gpt_forward(params, seq, ocfg, M.causal_mask(n))[0]).reshape(B, n, V)[:, -1]
        assert rel_err(logits, want) < TOL, seq.shape
        assert np.array_equal(np.argmax(logits, -1), np.argmax(want, -1))
        full = gen.full_logits(seq)
        assert rel_err(logits, full) < TOL, seq.shape
    assert gen.len == T


def test_generate_slides_like_the_reference(nt):
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.generate import KVGenerator
    B, T, V = 1, 10, 7
    layers = trainer.build_gpt_layers(B, T, 32, 2, 1, V, adam=False, seed=3)
    gen = KVGenerator(layers)
    out = gen.generate(np.array([[1, 2, 3]]), 12)
    assert out.shape == (1, 15)
    # replay with the reference's procedure: full window every time, sliding once it is full
    seq = [1, 2, 3]
    for _ in range(12):
        seq.append(int(np.argmax(gen.full_logits(np.array([seq[-T:]])), -1)[0]))
    assert seq == list(out[0])


def test_graph_replay_and_eager_steps_agree(nt):
    """The decode step is captured once (position in a device scalar) and replayed; the greedy continuation picks its
    tokens on the device.  Both must reproduce the eager, host-driven loop token for token, across two prompts."""
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.generate import KVGenerator
    B, T, V = 2, 32, 29
    layers = trainer.build_gpt_layers(B, T, 64, 2, 2, V, adam=False, seed=5)
    eager, graph = KVGenerator(layers, use_graph=False), KVGenerator(layers, use_graph=True)
    for prompt in (np.array([[3, 1, 4, 1], [5, 9, 2, 6]]), np.array([[7, 7], [0, 28]])):
        a = eager.generate(prompt, 20)
        b = graph.generate(prompt, 20)
        assert np.array_equal(a, b)
        assert graph.graph is not None and 0 < graph.graph_kernels < 80
    # sampled path (host picks the token): one graph launch per token, same logits as the eager step
    l0, l1 = eager.prefill(prompt), graph.prefill(prompt)
    assert rel_err(l1, l0) < TOL
    rs = np.random.RandomState(1)
    for _ in range(10):
        tok = rs.randint(0, V, B)
        l0, l1 = eager.step(tok), graph.step(tok)
        assert rel_err(l1, l0) < 1e-6
    # a window that fills up mid-way: cached steps, then the reference's sliding recompute
    assert np.array_equal(eager.generate(prompt, 40), graph.generate(prompt, 40))
    graph.close()
