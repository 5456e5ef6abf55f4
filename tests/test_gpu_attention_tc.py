"""The tcgen05 / TMEM attention kernels (csrc/attention_tc.cu, opt-in with NTB200_ATT_TC=1) for dh = 64, 16 <= N <= 64,
through the same C ABI as the default kernels (nt_attention_fwd / bwd [_split]) against the oracle (attention.py:31-46,
63-84).  The engine is chosen per process, so the test runs itself in a child process with the switch set."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r + '/tests')
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import ops
from oracle import numpy_ref as R
from oracle import models as M
from conftest import rel_err
nt.init(0)
worst = 0.0
for B, N, H, causal in [(3, 49, 6, False), (2, 64, 1, True), (5, 17, 3, False), (300, 49, 2, False), (2, 33, 5, True)]:
    rs = np.random.RandomState(B + N)
    qkv = rs.randn(B, N, 3 * H * 64)
    res = rs.randn(B, N, H * 64)
    dout = rs.randn(B, N, H * 64)
    mask = M.causal_mask(N) if causal else None
    O, P, _ = R.attention_fwd(qkv, H, mask)
    dq = R.attention_bwd(dout, qkv, P, H)
    kind = ops.MASK_CAUSAL if causal else ops.MASK_NONE
    dqkv_d = nt.as_device(qkv)
    out, Pd, _ = ops.attention_fwd(dqkv_d, H, None, kind, residual=nt.as_device(res), split_out=True)
    e = [rel_err(out, O + res), rel_err(Pd, P)]
    sp = getattr(out, "_split", None)               # bf16 hi / lo planes for the next Linear (only emitted for >= 2048 rows)
    if sp is not None:
        bf = lambda a: (a.numpy().view(np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(-1)[:B * N * H * 64]
        e.append(rel_err((bf(sp[0]) + bf(sp[1])).reshape(B, N, H * 64), O + res))
    out2, P2, _ = ops.attention_fwd(dqkv_d, H, None, kind)
    g = ops.attention_bwd(nt.as_device(dout), dqkv_d, P2, H, kind, split_out=True, att_out=out2)
    e.append(rel_err(g, dq))
    worst = max(worst, max(e))
    assert max(e) < 1e-4, (B, N, H, causal, e)
print("ATT_TC_OK worst %%.2e" %% worst)
"""


def test_attention_tc_kernels_vs_oracle():
    env = dict(os.environ, NTB200_ATT_TC="1")
    cp = subprocess.run([sys.executable, "-c", CHILD % (ROOT, ROOT)], env=env, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert cp.returncode == 0 and "ATT_TC_OK" in cp.stdout, (cp.stdout + cp.stderr)[-3000:]
    # the kernels really ran: the library reports the engine through its SASS (checked on the CPU side) and the switch
    assert "worst" in cp.stdout


def test_attention_tcl_forward_vs_oracle():
    """The tcgen05 forward for 64 < N <= 256 (csrc/attention_tcl.cu, the default for the GPT shapes): output (+ residual),
    the probabilities P incl. exact zeros above the causal diagonal, and the bf16 hi / lo planes, against the oracle; ragged
    N (not a multiple of 128 / 64 / 4), more items than CTAs, and the backward of attention_bf16.cu fed with its P."""
    import numpy_transformer_b200 as nt
    from numpy_transformer_b200 import ops
    from oracle import numpy_ref as R
    from oracle import models as M
    from conftest import rel_err
    nt.init(0)
    nt.set_gemm_mode(1)
    assert os.environ.get("NTB200_ATT_TCL", "1") != "0"
    worst = 0.0
    for B, N, H, causal in [(2, 256, 2, True), (1, 200, 3, True), (2, 128, 1, False), (3, 192, 2, False), (1, 100, 2, True),
                            (2, 250, 1, False), (1, 65, 1, True), (40, 256, 6, True)]:
        rs = np.random.RandomState(B + N)
        qkv = rs.randn(B, N, 3 * H * 64)
        res = rs.randn(B, N, H * 64)
        dout = rs.randn(B, N, H * 64)
        mask = M.causal_mask(N) if causal else None
        O, P, _ = R.attention_fwd(qkv, H, mask)
        kind = ops.MASK_CAUSAL if causal else ops.MASK_NONE
        qd = nt.as_device(qkv)
        out, Pd, _ = ops.attention_fwd(qd, H, None, kind, residual=nt.as_device(res), split_out=True)
        Ph = np.asarray(Pd)
        e = [rel_err(out, O + res), rel_err(Ph, P)]
        if causal:
            assert not np.triu(Ph, 1).any(), "probabilities above the causal diagonal must be exact zeros"
        sp = getattr(out, "_split", None)
        bf = lambda a: (a.numpy().view(np.uint16).astype(np.uint32) << 16).view(np.float32).reshape(-1)
        if sp is not None:
            e.append(rel_err((bf(sp[0]) + bf(sp[1]))[:B * N * H * 64].reshape(B, N, H * 64), O + res))
        out2, P2, _ = ops.attention_fwd(qd, H, None, kind)
        g = ops.attention_bwd(nt.as_device(dout), qd, P2, H, kind, att_out=out2)
        e.append(rel_err(g, R.attention_bwd(dout, qkv, P, H)))
        # the TMA operand path: qkv also given as the bf16 planes the QKV Linear's epilogue writes
        hi, lo, _ = ops.split_bf16(qd, B * N, 3 * H * 64)
        qd._split = (hi, lo)
        out3, P3, _ = ops.attention_fwd(qd, H, None, kind, residual=nt.as_device(res), split_out=True)
        assert qd._split is None, "the planes path was not taken"
        e += [rel_err(out3, O + res), rel_err(P3, P)]
        if causal:
            assert not np.triu(np.asarray(P3), 1).any()
        # the tcgen05 backward: dout and qkv as planes, P recomputed from the forward's row log-sum-exp
        qd._split = (hi, lo)
        out4, P4, _ = ops.attention_fwd(qd, H, None, kind)                     # no residual: att_out for the row sums
        dd = nt.as_device(dout)
        dd._split = ops.split_bf16(dd, B * N, H * 64)[:2]
        g2 = ops.attention_bwd(dd, qd, P4, H, kind, split_out=True, att_out=out4)
        assert ops._last["attention_bwd"] == "tcgen05-planes"
        e.append(rel_err(g2, R.attention_bwd(dout, qkv, P, H)))
        sp = getattr(g2, "_split", None)
        if sp is not None:
            e.append(rel_err((bf(sp[0]) + bf(sp[1]))[:B * N * 3 * H * 64].reshape(B, N, 3 * H * 64), R.attention_bwd(dout, qkv, P, H)))
        worst = max(worst, max(e))
        assert max(e) < 1e-4, (B, N, H, causal, e)
    print("attention_tcl worst rel err %.2e" % worst)


def test_attention_tcl_shape_sweep_vs_oracle():
    """Forward (TMA operand path) + backward of csrc/attention_tcl.cu over ragged sequence lengths, head counts and batch
    sizes that change the work-item walk (one / two query tiles, one / two key blocks, 1 ... 3 items per CTA, odd pair counts),
    causal and not, against the oracle.  Two consecutive calls per shape: the second one reuses barriers / TMEM of a warm CTA
    state only through fresh launches, but catches anything left behind in global scratch."""
    import numpy_transformer_b200 as nt
    from numpy_transformer_b200 import ops
    from oracle import numpy_ref as R
    from oracle import models as M
    from conftest import rel_err
    nt.init(0)
    nt.set_gemm_mode(1)
    rs = np.random.RandomState(7)
    shapes = [(1, 96, 1), (3, 127, 2), (2, 129, 3), (5, 160, 1), (1, 255, 4), (7, 256, 1), (150, 72, 1), (37, 256, 4),
              (75, 130, 2), (2, 68, 6)]
    worst = 0.0
    for B, N, H in shapes:
        for causal in (True, False):
            qkv = rs.randn(B, N, 3 * H * 64) * rs.choice([0.3, 1.0, 2.0])
            dout = rs.randn(B, N, H * 64)
            mask = M.causal_mask(N) if causal else None
            O, P, _ = R.attention_fwd(qkv, H, mask)
            ref_g = R.attention_bwd(dout, qkv, P, H)
            kind = ops.MASK_CAUSAL if causal else ops.MASK_NONE
            for rep in range(2):
                qd, dd = nt.as_device(qkv), nt.as_device(dout)
                qd._split = ops.split_bf16(qd, B * N, 3 * H * 64)[:2]
                dd._split = ops.split_bf16(dd, B * N, H * 64)[:2]
                out, Pd, _ = ops.attention_fwd(qd, H, None, kind)
                g = ops.attention_bwd(dd, qd, Pd, H, kind, att_out=out)
                assert ops._last["attention_bwd"] == "tcgen05-planes"
                e = [rel_err(out, O), rel_err(Pd, P), rel_err(g, ref_g)]
                worst = max(worst, max(e))
                assert max(e) < 1e-4, (B, N, H, causal, rep, e)
    print("attention_tcl sweep worst rel err %.2e" % worst)
