"""The three GEMM engines behind nt_linear_* (include/ntb200.h nt_set_gemm_mode) against fp64 NumPy:
  0 fp32 SIMT            tolerance 1e-4   (exact fp32)
  1 tcgen05 3xTF32       tolerance 1e-4   (default; the fp32/TF32 path of the spec)
  2 tcgen05 1xTF32       tolerance 1e-2   (fast path; measures ~1e-3)
Shapes cover every Linear of the README ViT, ragged edges, unaligned pitches (SIMT fallback inside
modes 1/2) and the long split-K weight-gradient contraction."""
import numpy as np
import pytest
from conftest import rel_err

pytestmark = pytest.mark.gpu
TOLS = {0: 1e-4, 1: 1e-4, 2: 1e-2}
SHAPES = [(4900, 16, 96), (4900, 96, 288), (4900, 96, 192), (4900, 192, 96), (100, 4704, 96), (100, 96, 10),
          (128, 32, 128), (1, 8, 8), (77, 40, 200), (513, 100, 36), (300, 384, 1152), (64, 50, 30), (2, 960, 192)]


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    yield nt
    nt.set_gemm_mode(1)


@pytest.mark.parametrize("mode", [0, 1, 2])
@pytest.mark.parametrize("M,K,N", SHAPES)
def test_linear_three_gemms(nt, mode, M, K, N):
    from numpy_transformer_b200 import ops, DeviceArray
    nt.set_gemm_mode(mode)
    tol = TOLS[mode]
    rs = np.random.RandomState(M * 31 + K * 7 + N)
    x, w, b, dy = rs.randn(M, K), rs.randn(K, N) / np.sqrt(K), rs.randn(N), rs.randn(M, N)
    res = rs.randn(M, N)
    dx_, dw_, db_, ddy, dres = (DeviceArray.from_host(a) for a in (x, w, b, dy, res))
    y, pre = ops.linear_fwd(dx_, dw_, db_, ops.ACT_GELU, want_pre=True, residual=dres)
    from oracle import numpy_ref as R
    ref_pre = x @ w + b
    assert rel_err(pre, ref_pre) < tol
    assert rel_err(y, R.gelu_fwd(ref_pre) + res) < tol
    dx = ops.linear_bwd_input(ddy, dw_)
    assert rel_err(dx, dy @ w.T) < tol
    pre_k = DeviceArray.from_host(rs.randn(M, K))
    add_k = DeviceArray.from_host(rs.randn(M, K))
    dx2 = ops.linear_bwd_input(ddy, dw_, ops.ACT_RELU, pre_k, add_k)
    assert rel_err(dx2, (dy @ w.T) * (np.asarray(pre_k) > 0) + np.asarray(add_k)) < tol
    g, gb = DeviceArray.zeros((K, N)), DeviceArray.zeros((N,))
    ops.linear_bwd_params(dx_, ddy, g, gb)
    ops.linear_bwd_params(dx_, ddy, g, gb)               # accumulates (+=)
    assert rel_err(g, 2 * (x.T @ dy)) < tol
    assert rel_err(gb, 2 * dy.sum(0)) < 1e-4


def test_mode_is_validated(nt):
    with pytest.raises(nt.NtError):
        nt.set_gemm_mode(7)
