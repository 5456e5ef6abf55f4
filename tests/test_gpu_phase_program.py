"""Phase programs on the B200 (csrc/phase_prog.cu, `nt_fuse_begin / nt_fuse_end`): the launch-bound GPT shapes — the
gpt_character training step (BASELINE.json configs[2]) and one KV-cache decode step — as ONE cluster kernel each.

The same device source is checked against the oracle on the CPU (tests/test_phase_program_cpu.py, SIMT emulator); here the
real kernel runs through the C ABI and is held (a) to the per-op kernels it replaces, (b) to the oracle, and (c) to its
launch budget (SURVEY.md 8(b) `nt_gpt_block_fwd/bwd`: <= 6 launches per gpt_character step, <= 4 per decoded token)."""
import ctypes

import numpy as np
import pytest
from conftest import rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    nt.set_gemm_mode(1)
    return nt


def fuse_stats():
    from numpy_transformer_b200._lib import lib
    a, b, c, d = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_int(0)
    lib.nt_fuse_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
    return a.value, b.value, c.value, d.value


@pytest.mark.parametrize("shape", [(2, 5, 192, 3, 1, 26), (2, 8, 16, 2, 2, 11), (1, 16, 64, 2, 2, 50)])
def test_fused_gpt_step_vs_per_op_kernels_and_oracle(nt, shape, monkeypatch):
    """Four SGD steps (eager, captured, two replays) of a <= 16-row GPT: the program path against the per-op kernels it
    replaces (loss and every parameter, 1e-5) and against the oracle (1e-4); launch budget of the captured step."""
    from numpy_transformer_b200 import trainer
    from oracle import models as M
    B, T, D, H, L, V = shape
    cfg = M.GPTConfig(batch=B, context=T, embed_dim=D, heads=H, blocks=L, vocab=V)
    toks, labs = M.synthetic_gpt_batch(cfg, seed=1)
    lr = 0.05
    runs = {}
    for fused in (False, True):
        monkeypatch.setenv("NTB200_FUSE", "1" if fused else "0")
        layers = trainer.build_gpt_layers(B, T, D, H, L, V, adam=False, seed=0)
        ts = trainer.TrainStep(layers, "gpt", (B, T), lr, adam=False)
        assert ts._fuse == fused                               # <= 16 token rows: NTB200_FUSE=1 selects the program path
        before = fuse_stats()
        losses = [ts.step(toks, labs) for _ in range(4)]
        after = fuse_stats()
        leaves = [np.asarray(x, dtype=np.float64).copy() for x in M.tree_leaves(trainer.save_walk(layers))]
        runs[fused] = (losses, leaves, ts.graph_kernels, [y - x for x, y in zip(before[:3], after[:3])])
        ts.close()
    (l0, p0, k0, s0), (l1, p1, k1, s1) = runs[False], runs[True]
    assert s0 == [0, 0, 0]                                     # per-op path: nothing recorded
    launches, phases, barriers = s1
    assert launches >= 3 and phases >= 20 * launches and barriers < phases, s1
    assert k1 <= 6 < k0, (k1, k0)
    assert np.allclose(l0, l1, rtol=1e-5, atol=0), (l0, l1)
    for i, (a, b) in enumerate(zip(p1, p0)):
        assert rel_err(a, b) < 1e-5, i
    params = M.gpt_init(cfg, 0)
    for i in range(4):
        ref = M.gpt_step(params, None, toks, labs, lr, cfg, adam=False)
        assert abs(l1[i] - ref["loss"]) < 1e-4 * abs(ref["loss"]), (i, l1[i], ref["loss"])
        params = ref["params"]
    for i, (a, b) in enumerate(zip(p1, M.tree_leaves(params))):
        assert rel_err(a.reshape(-1), np.asarray(b).reshape(-1)) < 1e-4, i


def test_fused_decode_step_vs_per_op_kernels(nt, monkeypatch):
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.generate import KVGenerator
    B, T, D, H, L, V = 2, 24, 64, 2, 2, 50
    rs = np.random.RandomState(0)
    prompt = rs.randint(0, V, (B, 5))
    out = {}
    for fused in ("0", "1"):
        monkeypatch.setenv("NTB200_FUSE", fused)
        layers = trainer.build_gpt_layers(B, T, D, H, L, V, adam=False, seed=4)
        gen = KVGenerator(layers)
        logits = gen.prefill(prompt)
        seq, per_step = [], []
        for _ in range(8):                                     # eager, captured, replays
            nxt = np.argmax(logits, -1)
            logits = gen.step(nxt)
            seq.append(nxt)
            per_step.append(logits.copy())
        toks = gen.greedy_steps(np.argmax(logits, -1), 6)      # device-side greedy continuation
        out[fused] = (np.array(per_step), toks, gen.graph_kernels)
        gen.close()
    (la, ta, ka), (lb, tb, kb) = out["0"], out["1"]
    assert rel_err(lb, la) < 1e-5
    assert np.array_equal(ta, tb)
    assert kb <= 4 < ka, (kb, ka)


def test_program_keeps_stream_order_around_ops_it_cannot_express(nt):
    """A fuse region with a 17-row Linear in the middle: the pending phases are launched, the big op runs its own kernel,
    recording resumes — same numbers as without the region."""
    from numpy_transformer_b200 import ops
    from numpy_transformer_b200._lib import lib
    from numpy_transformer_b200.device import DeviceArray
    rs = np.random.RandomState(3)
    w1, w2, w3 = (DeviceArray.from_host(rs.randn(32, 32).astype(np.float32) / 6) for _ in range(3))
    x4 = DeviceArray.from_host(rs.randn(4, 32).astype(np.float32))
    x17 = DeviceArray.from_host(rs.randn(17, 32).astype(np.float32))

    def chain():
        a = ops.linear_fwd(x4, w1, act=ops.ACT_RELU)
        b = ops.linear_fwd(x17, w2)                            # not expressible: more than 16 rows
        c = ops.linear_fwd(a, w3, residual=x4)
        return np.asarray(a).copy(), np.asarray(b).copy(), np.asarray(c).copy()

    want = chain()
    before = fuse_stats()
    lib.nt_fuse_begin()
    a = ops.linear_fwd(x4, w1, act=ops.ACT_RELU)
    b = ops.linear_fwd(x17, w2)
    c = ops.linear_fwd(a, w3, residual=x4)
    lib.nt_fuse_end()
    got = (np.asarray(a), np.asarray(b), np.asarray(c))
    after = fuse_stats()
    assert after[0] - before[0] == 2 and after[1] - before[1] == 2
    for g, w in zip(got, want):
        assert rel_err(g, w) < 1e-5
    assert after[3] >= 1                                       # CTAs per cluster the device granted
