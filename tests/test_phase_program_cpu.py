"""The phase-program kernel (csrc/phase_kernel.cuh + phase_record.h) on the CPU.

The device source is compiled UNCHANGED with g++ under the SIMT emulator of tests/simt_emu (a fibre per CUDA thread, real CTA
/ warp / cluster barriers, shuffles) and driven with NumPy arrays as device memory.  The op sequences below are the ones
`TrainStep` / `KVGenerator` issue for a GPT layer list (trainer.py, generate.py, dropin/gpt/attdecoderblock.py); results are held
to the oracle (oracle/models.py, fp64) at the 1e-4 bar of the GPU suite, for several cluster sizes, both fibre scheduling
orders (a missing barrier shows up in one of them) and with the dependency-tracked barriers against a barrier after
every phase (must be bit-identical).  No GPU involved; the GPU suite runs the same programs through the C ABI."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import models as OM

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "simt_emu")
F32, I32 = np.float32, np.int32


def _build():
    out_dir = os.path.join(EMU_DIR, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libphase_emu.so")
    srcs = [os.path.join(EMU_DIR, "phase_emu.cpp"), os.path.join(EMU_DIR, "simt_emu.h")]
    csrc = os.path.join(os.path.dirname(HERE), "numpy_transformer_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in ("phase_kernel.cuh", "phase_record.h", "phase_prog.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        # -Dntph=... / -Bsymbolic: libntb200.so (loaded RTLD_GLOBAL by other tests of the session) exports nvcc's host stubs of
        # the same device functions; the emulator build must bind to its own definitions, not to those
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-Dntph=ntph_emu",
                               "-Wl,-Bsymbolic", "-o", so, srcs[0]])
    return ctypes.CDLL(so)


@pytest.fixture(scope="module")
def emu():
    import shutil
    if shutil.which("g++") is None:
        pytest.skip("no g++ here: the emulator build of the phase kernel needs a host C++ compiler")
    return Emu(_build())


class Emu:
    """ctypes face of libphase_emu.so; `a(...)` allocates "device" arrays that stay alive until reset()."""

    def __init__(self, dll):
        self.dll = dll
        self.keep = []

    def reset(self):
        self.keep = []
        self.dll.emu_begin()

    def a(self, src=None, shape=None, dtype=F32):
        arr = np.zeros(shape, dtype) if src is None else np.ascontiguousarray(np.asarray(src).astype(dtype))
        self.keep.append(arr)
        return arr

    @staticmethod
    def _arg(v):
        if v is None:
            return ctypes.c_void_p(None)
        if isinstance(v, np.ndarray):
            return ctypes.c_void_p(v.ctypes.data)
        if isinstance(v, float):
            return ctypes.c_float(v)
        return ctypes.c_int(int(v))

    def call(self, name, *args):
        rc = getattr(self.dll, "emu_" + name)(*[self._arg(v) for v in args])
        return rc

    def op(self, name, *args):
        assert self.call(name, *args) == 0, "%s was refused by the recorder" % name

    def run(self, nctas=4, reverse=0, all_barriers=0):
        assert self.dll.emu_run(int(nctas), int(reverse), int(all_barriers)) == 0

    def counts(self):
        return self.dll.emu_phase_count(), self.dll.emu_barrier_count()


def rel(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


# ------------------------------------------------------------------------------------------------ the GPT step as a program
def record_gpt_step(E, params, tokens, labels, cfg):
    """Record forward + loss + backward of trainer.TrainStep for a GPT layer list (same op order as the layer objects);
    -> dict of the arrays that come back (loss, logits, argmax, gradient tree like oracle.models.gpt_backward)."""
    B, T = tokens.shape
    D, H, V = cfg.embed_dim, cfg.heads, cfg.vocab
    M, dh = B * T, cfg.embed_dim // cfg.heads
    P = [E.a(p) for p in OM.tree_leaves(params)]
    G = [E.a(shape=p.shape) for p in P]
    it = iter(range(len(P)))
    etok_i, epos_i = next(it), next(it)
    blocks = []
    for _ in range(cfg.blocks):
        blocks.append({k: (next(it), next(it)) for k in ("ln", "qkv", "ln1", "fc0", "fc1", "fco")})
    lnf = (next(it), next(it))
    c0, c1 = (next(it), next(it)), (next(it), next(it))
    tok = E.a(tokens.reshape(-1), dtype=I32)
    lab = E.a(labels.reshape(-1), dtype=I32)

    def linear(x, wb, act=0, want_pre=False, residual=None):
        w, b = P[wb[0]], P[wb[1]]
        y = E.a(shape=(M, w.shape[1]))
        pre = E.a(shape=(M, w.shape[1])) if want_pre else None
        E.op("linear_fwd", x, w, b, y, M, w.shape[0], w.shape[1], act, pre, residual)
        return (y, pre) if want_pre else y

    def layernorm(x, gb):
        y, mean, rstd = E.a(shape=(M, D)), E.a(shape=(M,)), E.a(shape=(M,))
        E.op("layernorm_fwd", x, P[gb[0]], P[gb[1]], y, mean, rstd, M, D, 1e-5)
        return y, (x, mean, rstd)

    def linear_bwd(dy, x, wb, act=0, pre=None, addend=None):
        w = P[wb[0]]
        E.op("linear_bwd_params", x, dy, G[wb[0]], G[wb[1]], M, w.shape[0], w.shape[1])
        dx = E.a(shape=(M, w.shape[0]))
        E.op("linear_bwd_input", dy, w, dx, M, w.shape[0], w.shape[1], act, pre, addend)
        return dx

    def layernorm_bwd(dy, saved, gb, addend=None):
        x, mean, rstd = saved
        dx = E.a(shape=(M, D))
        E.op("layernorm_bwd", dy, x, mean, rstd, P[gb[0]], dx, G[gb[0]], G[gb[1]], M, D, addend)
        return dx

    x = E.a(shape=(M, D))
    E.op("embedding_fwd", tok, P[etok_i], P[epos_i], x, B, T, D, V)
    saved = []
    for blk in blocks:
        s = {"x": x}
        s["out0"], s["ln"] = layernorm(x, blk["ln"])
        s["out2"], s["pre2"] = linear(s["out0"], blk["fc0"], act=1, want_pre=True)
        s["out6"] = linear(s["out2"], blk["fc1"], residual=x)
        s["outnorm"], s["ln1"] = layernorm(s["out6"], blk["ln1"])
        s["qkv"] = linear(s["outnorm"], blk["qkv"])
        s["rkk"], s["P"] = E.a(shape=(M, D)), E.a(shape=(B, H, T, T))
        E.op("attention_fwd", s["qkv"], 1, s["rkk"], s["P"], B, T, H, dh, None)
        x = linear(s["rkk"], blk["fco"], residual=s["out6"])
        saved.append(s)
    z, lnf_saved = layernorm(x, lnf)
    h = linear(z, c0)
    logits = linear(h, c1)
    loss, row_loss = E.a(shape=(1,)), E.a(shape=(M,))
    grad, prob, amax = E.a(shape=(M, V)), E.a(shape=(M, V)), E.a(shape=(M,), dtype=I32)
    E.op("cross_entropy", logits, lab, None, float(M), loss, row_loss, grad, prob, amax, M, V)
    d = linear_bwd(grad, h, c1)
    d = linear_bwd(d, z, c0)
    d = layernorm_bwd(d, lnf_saved, lnf)
    for blk, s in zip(reversed(blocks), reversed(saved)):
        delta = d
        d0 = linear_bwd(delta, s["rkk"], blk["fco"])
        dqkv = E.a(shape=(M, 3 * D))
        E.op("attention_bwd", d0, s["qkv"], s["P"], dqkv, B, T, H, dh)
        dd = linear_bwd(dqkv, s["outnorm"], blk["qkv"])
        qkvdelta = layernorm_bwd(dd, s["ln1"], blk["ln1"], addend=delta)
        dd = linear_bwd(qkvdelta, s["out2"], blk["fc1"], act=1, pre=s["pre2"])
        dd = linear_bwd(dd, s["out0"], blk["fc0"])
        d = layernorm_bwd(dd, s["ln"], blk["ln"], addend=qkvdelta)
    E.op("embedding_bwd", tok, d, G[etok_i], G[epos_i], B, T, D, V)
    return dict(loss=loss, logits=logits, argmax=amax, prob=prob, G=G, dx0=d)


def check_gpt_step(E, cfg, nctas, reverse, all_barriers=0, seed=0, tol=1e-4):
    params = OM.gpt_init(cfg, seed=seed)
    tokens, labels = OM.synthetic_gpt_batch(cfg, seed=seed + 1)
    tokens[0, 1] = tokens[0, 0]                                  # a repeated token: the embedding scatter must accumulate
    ref = OM.gpt_step(params, OM.adam_state(params), tokens, labels, 1e-3, cfg)
    E.reset()
    out = record_gpt_step(E, params, tokens, labels, cfg)
    phases, barriers = E.counts()
    E.run(nctas, reverse, all_barriers)
    assert rel(out["loss"][0], ref["loss"]) < tol
    assert rel(out["logits"], ref["logits"].reshape(out["logits"].shape)) < tol
    np.testing.assert_array_equal(out["argmax"], np.argmax(out["prob"], axis=-1))
    np.testing.assert_array_equal(out["argmax"], ref["argmax"].reshape(-1))
    worst = 0.0
    for g, r in zip(out["G"], OM.tree_leaves(ref["grads"])):
        worst = max(worst, rel(g, np.asarray(r).reshape(g.shape)))
    assert worst < tol, worst
    return out, phases, barriers


TINY = OM.GPTConfig(batch=2, context=5, embed_dim=32, heads=2, blocks=2, vocab=26)


@pytest.mark.parametrize("nctas,reverse", [(1, 0), (3, 1), (4, 0), (4, 1)])
def test_gpt_step_program_vs_oracle(emu, nctas, reverse):
    out, phases, barriers = check_gpt_step(emu, TINY, nctas, reverse)
    assert phases == 1 + 2 * 7 + 3 + 2 + (2 * 2 + 1) + 2 * 11 + 1      # embed, blocks, head, loss; head, blocks, embed backward
    assert barriers < phases - 10            # weight-gradient phases and the loss scalar share barrier intervals


def test_tracked_barriers_equal_a_barrier_after_every_phase(emu):
    a, _, _ = check_gpt_step(emu, TINY, 4, 0, all_barriers=0)
    a = [g.copy() for g in a["G"]] + [a["logits"].copy()]
    b, _, _ = check_gpt_step(emu, TINY, 4, 0, all_barriers=1)
    for x, y in zip(a, b["G"] + [b["logits"]]):
        np.testing.assert_array_equal(x, y)


def test_emulator_notices_missing_cluster_barriers(emu):
    """Self-test of the test vehicle: with every cluster barrier stripped from the program the CTAs run ahead of each other
    and the result is wrong in at least one scheduling order — i.e. the comparisons above would catch a barrier the
    dependency tracking forgot."""
    wrong = 0
    for reverse in (0, 1):
        try:
            check_gpt_step(emu, TINY, 4, reverse, all_barriers=2)
        except AssertionError:
            wrong += 1
    assert wrong >= 1


@pytest.mark.parametrize("nctas", [4, 16])
def test_gpt_character_shape(emu, nctas):
    """BASELINE.json configs[2] at its real shape: 2 x 5 tokens, D = 192, 3 heads (dh = 64), one block, 26 characters — on 4
    CTAs and on the 16-CTA cluster the B200 launch uses (8192 fibres)."""
    out, phases, barriers = check_gpt_step(emu, OM.GPT_CHAR, nctas, 1)
    assert (phases, barriers) == (30, 22)


def test_sixteen_rows_wide_vocabulary(emu):
    """16 token rows (the row-tile limit), T = 8, a head wider than one pass of 32 column quads per CTA."""
    check_gpt_step(emu, OM.GPTConfig(batch=2, context=8, embed_dim=32, heads=2, blocks=1, vocab=301), 2, 0)


# ------------------------------------------------------------------------------------------------ one decode step
def test_decode_steps_vs_oracle(emu):
    """generate.KVGenerator._enqueue_step as a program: every step's logits against oracle.models.gpt_forward on the prefix,
    greedy token chosen on the "device", position counter advanced by the program itself."""
    cfg = OM.GPTConfig(batch=2, context=12, embed_dim=32, heads=2, blocks=2, vocab=26)
    E = emu
    params = OM.gpt_init(cfg, seed=3)
    B, D, H, V, Tmax = cfg.batch, cfg.embed_dim, cfg.heads, cfg.vocab, cfg.context
    dh = D // H
    rs = np.random.RandomState(5)
    prompt = rs.randint(0, V, (B, 4))
    E.reset()
    P = [E.a(p) for p in OM.tree_leaves(params)]
    it = iter(range(len(P)))
    etok_i, epos_i = next(it), next(it)
    blocks = [{k: (next(it), next(it)) for k in ("ln", "qkv", "ln1", "fc0", "fc1", "fco")} for _ in range(cfg.blocks)]
    lnf = (next(it), next(it))
    c0, c1 = (next(it), next(it)), (next(it), next(it))
    # prefill by the oracle: caches hold k / v of the prompt positions
    T0 = prompt.shape[1]
    _, acts = OM.gpt_forward(params, prompt, cfg, OM.causal_mask(T0))
    kc = [E.a(shape=(B, Tmax, D)) for _ in blocks]
    vc = [E.a(shape=(B, Tmax, D)) for _ in blocks]
    for l in range(cfg.blocks):
        qkv = acts["caches"][l]["qkv"]
        kc[l][:, :T0] = qkv[..., D:2 * D]
        vc[l][:, :T0] = qkv[..., 2 * D:]
    d_pos = E.a([T0], dtype=I32)
    d_tok = E.a(shape=(B,), dtype=I32)
    d_hist = E.a(shape=(B, Tmax + 1), dtype=I32)
    seq = prompt.copy()
    nxt = rs.randint(0, V, (B,))
    for step in range(4):
        d_tok[:] = nxt
        seq = np.concatenate([seq, nxt[:, None]], axis=1)
        x = E.a(shape=(B, D))
        E.op("decode_embed", d_tok, P[etok_i], P[epos_i], x, B, D, Tmax, d_pos, V)

        def ln_linear(xin, gb, wb, act=0):
            w = P[wb[0]]
            y = E.a(shape=(B, w.shape[1]))
            E.op("linear_ln_skinny", xin, P[gb[0]], P[gb[1]], w, P[wb[1]], y, B, w.shape[0], w.shape[1], act, None)
            return y

        def linear(xin, wb, residual=None):
            w = P[wb[0]]
            y = E.a(shape=(B, w.shape[1]))
            E.op("linear_fwd", xin, w, P[wb[1]], y, B, w.shape[0], w.shape[1], 0, None, residual)
            return y

        for l, blk in enumerate(blocks):
            hdn = ln_linear(x, blk["ln"], blk["fc0"], act=1)
            h = linear(hdn, blk["fc1"], residual=x)
            qkv = ln_linear(h, blk["ln1"], blk["qkv"])
            att = E.a(shape=(B, D))
            E.op("attention_decode_at", qkv, kc[l], vc[l], att, B, Tmax, H, dh, d_pos, 1)
            x = linear(att, blk["fco"], residual=h)
        z1 = ln_linear(x, lnf, c0)
        logits = linear(z1, c1)
        E.op("argmax_next", logits, V, V, B, d_tok, d_hist, Tmax + 1, d_pos)
        E.op("increment", d_pos)
        phases, barriers = E.counts()
        assert phases == 1 + 5 * cfg.blocks + 2 + 2
        E.run(3, step % 2)
        ref, _ = OM.gpt_forward(params, seq, cfg, OM.causal_mask(seq.shape[1]))
        assert rel(logits, ref[:, -1, :]) < 1e-4
        assert int(d_pos[0]) == T0 + step + 1
        np.testing.assert_array_equal(d_tok, np.argmax(logits, axis=-1))
        np.testing.assert_array_equal(d_hist[:, T0 + step + 1], d_tok)
        nxt = d_tok.copy()


@pytest.mark.parametrize("dh,append,pos", [(72, 1, 9), (72, 0, 9), (12, 1, 140), (64, 1, 0), (3, 0, 5)])
def test_decode_attention_phase(emu, dh, append, pos):
    """nt_attention_decode_at alone: head widths on both sides of the value prefetch (dh <= 64 / > 64, dh % 4 != 0), more cached
    rows than the prefetch covers, the very first position, append folded in or done beforehand."""
    E = emu
    rs = np.random.RandomState(dh * 7 + pos)
    B, H, Tmax = 2, 3, 150
    D = H * dh
    qkv = rs.randn(B, 3 * D)
    kc, vc = rs.randn(B, Tmax, D), rs.randn(B, Tmax, D)
    kc[:, pos + 1:], vc[:, pos + 1:] = 7e3, -7e3                   # beyond the window: must never be read
    want_k, want_v = kc.copy(), vc.copy()
    want_k[:, pos], want_v[:, pos] = qkv[:, D:2 * D], qkv[:, 2 * D:]
    if not append:
        kc, vc = want_k.copy(), want_v.copy()                          # nt_kv_append_at ran before
    E.reset()
    QKV, KC, VC, OUT, POS = E.a(qkv), E.a(kc), E.a(vc), E.a(shape=(B, D)), E.a([pos], dtype=I32)
    E.op("attention_decode_at", QKV, KC, VC, OUT, B, Tmax, H, dh, POS, append)
    E.run(4, pos % 2)
    q = qkv[:, :D].reshape(B, H, dh)
    k = want_k[:, :pos + 1].reshape(B, pos + 1, H, dh)
    v = want_v[:, :pos + 1].reshape(B, pos + 1, H, dh)
    s = np.einsum("bhd,bjhd->bhj", q, k) / np.sqrt(dh)
    p = np.exp(s - s.max(-1, keepdims=True))
    p /= p.sum(-1, keepdims=True)
    want = np.einsum("bhj,bjhd->bhd", p, v).reshape(B, D)
    assert rel(OUT, want) < 1e-5
    assert rel(KC, want_k) < 1e-7 and rel(VC, want_v) < 1e-7


# ------------------------------------------------------------------------------------------------ single ops, odd shapes
@pytest.mark.parametrize("M,K,N,act", [(1, 48, 7, 0), (3, 33, 130, 2), (10, 64, 36, 1), (16, 40, 5, 2), (7, 24, 1030, 0)])
def test_linear_phases_odd_shapes(emu, M, K, N, act):
    from oracle import numpy_ref as R
    E = emu
    rs = np.random.RandomState(M * 1000 + N)
    x, w, b, res = rs.randn(M, K), rs.randn(K, N) / np.sqrt(K), rs.randn(N), rs.randn(M, N)
    dy = rs.randn(M, N)
    E.reset()
    X, W, Bv, RES, DY = E.a(x), E.a(w), E.a(b), E.a(res), E.a(dy)
    if N % 4 == 0:
        W = E.a(np.concatenate([[0.0], w.reshape(-1)]))[1:].reshape(K, N)         # a weight panel off the 16-byte grid
    Y, PRE = E.a(shape=(M, N)), E.a(shape=(M, N))
    DX, DW, DB = E.a(shape=(M, K)), E.a(rs.randn(K, N)), E.a(rs.randn(N))
    dw0, db0 = DW.copy(), DB.copy()
    E.op("linear_fwd", X, W, Bv, Y, M, K, N, act, PRE, RES)
    E.op("linear_bwd_params", X, DY, DW, DB, M, K, N)
    E.op("linear_bwd_input", DY, W, DX, M, K, N, 0, None, X)
    assert E.counts() == (3, 0)                                   # three independent phases: no cluster barrier at all
    E.run(3, 1)
    pre = x @ w + b
    want = {0: pre, 1: np.maximum(pre, 0), 2: R.gelu_fwd(pre)}[act] + res
    assert rel(PRE, pre) < 1e-5 and rel(Y, want) < 1e-5
    assert rel(DW - dw0, x.T @ dy) < 1e-5 and rel(DB - db0, dy.sum(0)) < 1e-5       # accumulates into what was there
    assert rel(DX, dy @ w.T + x) < 1e-5


def test_activation_backward_epilogues(emu):
    from oracle import numpy_ref as R
    E = emu
    rs = np.random.RandomState(11)
    M, K, N = 5, 44, 28
    dy, w, pre = rs.randn(M, N), rs.randn(K, N), rs.randn(M, K)
    for act, f in ((1, lambda d: R.relu_bwd(d, pre)), (2, lambda d: R.gelu_bwd(d, pre))):
        E.reset()
        DY, W, PRE, DX = E.a(dy), E.a(w), E.a(pre), E.a(shape=(M, K))
        E.op("linear_bwd_input", DY, W, DX, M, K, N, act, PRE, None)
        E.run(2, 0)
        assert rel(DX, f(dy @ w.T)) < 1e-5


def test_attention_phases_no_mask_and_residual(emu):
    from oracle import numpy_ref as R
    E = emu
    rs = np.random.RandomState(2)
    B, N, H, dh = 2, 7, 3, 8
    D = H * dh
    qkv, res, dout = rs.randn(B, N, 3 * D), rs.randn(B, N, D), rs.randn(B, N, D)
    O, P, _ = R.attention_fwd(qkv, H, None)
    dqkv = R.attention_bwd(dout, qkv, P, H)
    E.reset()
    QKV, RES, DO = E.a(qkv), E.a(res), E.a(dout)
    OUT, PP, DQKV = E.a(shape=(B, N, D)), E.a(shape=(B, H, N, N)), E.a(shape=(B, N, 3 * D))
    E.op("attention_fwd", QKV, 0, OUT, PP, B, N, H, dh, RES)
    E.op("attention_bwd", DO, QKV, PP, DQKV, B, N, H, dh)
    assert E.counts() == (2, 1)
    E.run(4, 1)
    assert rel(OUT, O + res) < 1e-5 and rel(PP, np.asarray(P).reshape(PP.shape)) < 1e-5 and rel(DQKV, dqkv) < 1e-5


def test_recorder_refuses_what_a_program_cannot_express(emu):
    E = emu
    E.reset()
    x, w, y = E.a(shape=(17, 8)), E.a(shape=(8, 8)), E.a(shape=(17, 8))
    assert E.call("linear_fwd", x, w, None, y, 17, 8, 8, 0, None, None) != 0               # more than 16 rows
    assert E.call("linear_fwd", x, w, None, x, 4, 8, 8, 0, None, None) != 0                # in place: CTAs race on their slices
    big = E.a(shape=(1, 60000))
    assert E.call("linear_fwd", big, big, None, y, 1, 60000, 1, 0, None, None) != 0        # row panel beyond shared memory
    qkv = E.a(shape=(1, 40, 3 * 8))
    assert E.call("attention_fwd", qkv, 1, y, None, 1, 40, 1, 8, None) != 0                # N > 32 belongs to the big kernels
    assert E.counts() == (0, 0)


def test_label_out_of_range_sets_the_error_flag(emu):
    E = emu
    E.reset()
    logits = E.a(np.random.RandomState(0).randn(3, 6))
    lab = E.a([1, 9, 2], dtype=I32)
    loss, row_loss, grad, prob, amax = E.a(shape=(1,)), E.a(shape=(3,)), E.a(shape=(3, 6)), E.a(shape=(3, 6)), E.a(shape=(3,), dtype=I32)
    E.op("cross_entropy", logits, lab, None, 3.0, loss, row_loss, grad, prob, amax, 3, 6)
    E.run(2, 0)
    assert E.dll.emu_err_flag() == 2 and row_loss[1] == 0.0


# ------------------------------------------------------------------------------------------------ nt_gpt_block_fwd / _bwd
_BLOCK_FIELDS = ["ln_g", "ln_b", "w0", "b0", "w1", "b1", "ln1_g", "ln1_b", "wqkv", "bqkv", "wo", "bo",
                 "d_ln_g", "d_ln_b", "d_w0", "d_b0", "d_w1", "d_b1", "d_ln1_g", "d_ln1_b", "d_wqkv", "d_bqkv", "d_wo", "d_bo",
                 "out0", "mean0", "rstd0", "pre2", "out2", "out6", "outnorm", "mean1", "rstd1", "qkv", "P", "rkk"]


class NtGptBlock(ctypes.Structure):
    """Mirror of `struct NtGptBlock` (include/ntb200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in _BLOCK_FIELDS]


def test_header_struct_matches_the_mirror():
    import re
    src = open(os.path.join(os.path.dirname(HERE), "include", "ntb200.h")).read()
    body = re.search(r"typedef struct NtGptBlock \{(.*?)\} NtGptBlock;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\*\s*(\w+)", body)
    assert names == _BLOCK_FIELDS


@pytest.mark.parametrize("B,T,D,H,mask_kind", [(2, 5, 32, 2, 1), (1, 16, 48, 3, 0)])
def test_gpt_block_entry_points_vs_oracle(emu, B, T, D, H, mask_kind):
    """The op sequence behind nt_gpt_block_fwd / nt_gpt_block_bwd (csrc/gpt_block.h, shared by the library and the emulator build)
    against oracle.models.gpt_block_fwd / gpt_block_bwd: output, saved activations, input gradient and all twelve parameter
    gradients, accumulated on top of what the buffers held."""
    E = emu
    rs = np.random.RandomState(B * 100 + T)
    M, F = B * T, 4 * D

    def fc(i, o):
        return [rs.randn(i, o) / np.sqrt(i), rs.randn(o) * 0.1]
    blk = [[1 + 0.1 * rs.randn(D), 0.1 * rs.randn(D)], fc(D, 3 * D), [1 + 0.1 * rs.randn(D), 0.1 * rs.randn(D)], fc(D, F), fc(F, D), fc(D, D)]
    (g0, b0), (Wqkv, bqkv), (g1, b1), (W0, bb0), (W1, bb1), (Wo, bo) = blk
    x, dy = rs.randn(B, T, D), rs.randn(B, T, D)
    mask = OM.causal_mask(T) if mask_kind == 1 else None
    want_y, c = OM.gpt_block_fwd(x, blk, H, mask)
    want_dx, want_g = OM.gpt_block_bwd(dy, blk, c, H)
    E.reset()
    k = NtGptBlock()
    params = dict(ln_g=g0, ln_b=b0, w0=W0, b0=bb0, w1=W1, b1=bb1, ln1_g=g1, ln1_b=b1, wqkv=Wqkv, bqkv=bqkv, wo=Wo, bo=bo)
    held = {}
    for n, v in params.items():
        held[n] = E.a(v)
        held["d_" + n] = E.a(rs.randn(*np.shape(v)))                     # gradients accumulate on top of this
    start = {n: held["d_" + n].copy() for n in params}
    shapes = dict(out0=(M, D), mean0=(M,), rstd0=(M,), pre2=(M, F), out2=(M, F), out6=(M, D), outnorm=(M, D), mean1=(M,), rstd1=(M,),
                  qkv=(M, 3 * D), P=(B, H, T, T), rkk=(M, D))
    for n, shp in shapes.items():
        held[n] = E.a(shape=shp)
    for n in _BLOCK_FIELDS:
        setattr(k, n, held[n].ctypes.data)
    X, DY, Y, DX = E.a(x), E.a(dy), E.a(shape=(M, D)), E.a(shape=(M, D))
    E.dll.emu_gpt_block_bwd_workspace.restype = ctypes.c_size_t
    ws = E.a(shape=(E.dll.emu_gpt_block_bwd_workspace(B, T, D, H),))
    assert ws.size == 11 * M * D
    p = lambda a: ctypes.c_void_p(a.ctypes.data)
    assert E.dll.emu_gpt_block_fwd(ctypes.byref(k), p(X), p(Y), B, T, D, H, mask_kind) == 0
    assert E.counts()[0] == 7
    E.run(4, 0)
    assert rel(Y, want_y.reshape(M, D)) < 1e-5
    assert rel(held["qkv"], c["qkv"].reshape(M, 3 * D)) < 1e-5 and rel(held["P"], np.asarray(c["P"]).reshape(B, H, T, T)) < 1e-5
    assert rel(held["out6"], c["h"].reshape(M, D)) < 1e-5 and rel(held["pre2"], c["pre"].reshape(M, F)) < 1e-5
    assert E.dll.emu_gpt_block_bwd(ctypes.byref(k), p(X), p(DY), p(DX), p(ws), B, T, D, H, mask_kind) == 0
    assert E.counts()[0] == 11
    E.run(3, 1)
    assert rel(DX, want_dx.reshape(M, D)) < 1e-5
    order = [("ln_g", "ln_b"), ("wqkv", "bqkv"), ("ln1_g", "ln1_b"), ("w0", "b0"), ("w1", "b1"), ("wo", "bo")]
    for (wn, bn), (gw, gb) in zip(order, want_g):
        assert rel(held["d_" + wn] - start[wn], np.asarray(gw).reshape(start[wn].shape)) < 2e-5, wn
        assert rel(held["d_" + bn] - start[bn], np.asarray(gb).reshape(start[bn].shape)) < 2e-5, bn
