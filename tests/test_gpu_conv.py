"""GPU parity of the convolution path: net.Convolution.convolution_layer and
PatchEmbed.PatchEmbed_convolution of the drop-in mirror against vectors produced by the unmodified reference
(tests/golden/conv.npz, oracle/gen_golden.py:gen_conv) and against the oracle on larger shapes.
Tolerance: max|a-b|/max|b| <= 1e-4 per tensor (fp32-class path)."""
import numpy as np
import pytest
from conftest import load_golden, rel_err
from oracle import numpy_ref as R

pytestmark = pytest.mark.gpu
TOL = 1e-4
CASES = [(4, 0), (1, 1), (2, 1), ((2, 1), (0, 1)), (3, 0), (2, 0)]     # (stride, padding) of gen_golden.CONV_CASES


@pytest.fixture(scope="module")
def nt():
    import numpy_transformer_b200 as nt
    nt.init(0)
    nt.install()
    nt.set_gemm_mode(1)
    return nt


@pytest.mark.parametrize("i", range(len(CASES)))
def test_convolution_layer_golden(nt, i):
    from net.Convolution import convolution_layer
    g = load_golden("conv.npz")
    s, p = CASES[i]
    x, W, b = g["c%d_x" % i], g["c%d_W" % i], g["c%d_b" % i]
    conv = convolution_layer(W.shape[1], W.shape[0], list(W.shape[2:]), stride=s, padding=p, params=W, bias_params=b)
    assert rel_err(conv.params, W) < 1e-6 and conv.params.shape == W.shape
    y = conv.forward(x)
    assert tuple(conv.outshape) == g["c%d_y" % i].shape
    assert rel_err(y, g["c%d_y" % i]) < TOL
    dx = conv.backward(g["c%d_g" % i])
    assert rel_err(dx, g["c%d_dx" % i]) < TOL
    assert rel_err(conv.params_delta, g["c%d_dW" % i]) < TOL and rel_err(conv.bias_delta, g["c%d_db" % i]) < TOL
    conv.backward(g["c%d_g" % i])                                     # accumulate until setzero
    assert rel_err(conv.params_delta, g["c%d_dW2" % i]) < TOL
    conv.update(0.05)
    assert rel_err(conv.params, g["c%d_W1" % i]) < TOL and rel_err(conv.bias_params, g["c%d_b1" % i]) < TOL
    conv.setzero()
    assert not np.any(conv.params_delta) and not np.any(np.asarray(conv.bias_delta))
    # checkpoint round trip in the reference's (OC, C, KH, KW) layout
    saved = conv.save_model()
    other = convolution_layer(W.shape[1], W.shape[0], list(W.shape[2:]), stride=s, padding=p)
    other.restore_model(saved)
    assert rel_err(other.forward(x), conv.forward(x)) < 1e-6


def test_patch_embed_convolution_golden(nt):
    from PatchEmbed import PatchEmbed_convolution
    g = load_golden("conv.npz")
    np.random.seed(7)
    pe = PatchEmbed_convolution(12, (3, 2, 8, 8), 2, patchnorm=True)     # same seed as the generator: same draws
    assert rel_err(pe.convolution.params, g["pe_W"]) < 1e-6 and rel_err(pe.convolution.bias_params, g["pe_b"]) < 1e-6
    y = pe.forward(g["pe_x"])
    assert rel_err(y, g["pe_y"]) < TOL                                  # the PRE-LayerNorm tensor, as in the reference
    dx = pe.backward(g["pe_g"])
    assert rel_err(dx, g["pe_dx"]) < TOL
    assert rel_err(pe.convolution.params_delta, g["pe_dW"]) < TOL and rel_err(pe.convolution.bias_delta, g["pe_db"]) < TOL
    assert rel_err(pe.norm.gamma_delta, g["pe_dgamma"]) < TOL and rel_err(pe.norm.beta_delta, g["pe_dbeta"]) < TOL


@pytest.mark.parametrize("shape,oc,k,s,p", [((64, 3, 32, 32), 64, 3, 1, 1), ((100, 1, 28, 28), 96, 4, 4, 0),
                                            ((5, 4, 17, 13), 10, (5, 3), (2, 3), (2, 1)), ((0, 2, 8, 8), 4, 2, 2, 0)])
def test_convolution_vs_oracle(nt, shape, oc, k, s, p):
    """Larger / ragged shapes (the 64 x 3 x 32 x 32 case runs the BF16x3 tensor-core engine: 65536 rows), empty batch."""
    from net.Convolution import convolution_layer
    rs = np.random.RandomState(3)
    x = rs.randn(*shape)
    np.random.seed(5)
    conv = convolution_layer(shape[1], oc, k, stride=s, padding=p)
    W, b = conv.params, np.asarray(conv.bias_params)
    y = conv.forward(x)
    ref = R.conv_fwd(x, W, b, s, p)
    assert np.asarray(y).shape == ref.shape
    if shape[0] == 0:
        return
    assert rel_err(y, ref) < TOL
    gy = rs.randn(*ref.shape)
    dx = conv.backward(gy)
    rdx, rdW, rdb = R.conv_bwd(gy, x, W, s, p)
    assert rel_err(dx, rdx) < TOL and rel_err(conv.params_delta, rdW) < TOL and rel_err(conv.bias_delta, rdb) < TOL


def test_trainstep_with_conv_patch_embed(nt):
    """A ViT whose patch embedding is the convolution (transformer_of_image.py with patch_convolu=1, patchnorm off):
    kernel = stride = patch makes it the same function as PatchEmbed_flatten with the kernel tensor reshaped, so the
    two TrainSteps must follow the same trajectory."""
    from numpy_transformer_b200 import trainer
    from PatchEmbed import PatchEmbed_convolution, PatchEmbed_flatten
    B = 6
    rs = np.random.RandomState(0)
    images, labels = rs.rand(B, 1, 28, 28), rs.randint(0, 10, B)
    steps = []
    for kind in ("flatten", "conv"):
        layers = trainer.build_vit_layers(B, embed_dim=32, heads=2, blocks=2, seed=0)
        if kind == "flatten":
            flat = PatchEmbed_flatten(32, images.shape, 7, patchnorm=False)
            w0, b0 = np.asarray(flat.fullconnect.params), np.asarray(flat.fullconnect.bias_params)
            layers[0] = flat
        else:
            conv = PatchEmbed_convolution(32, images.shape, 7, patchnorm=False)
            conv.convolution.params = w0.T.reshape(32, 1, 4, 4)
            conv.convolution.bias_params = b0
            layers[0] = conv
        ts = trainer.TrainStep(layers, "vit", images.shape, lr=0.05)
        steps.append([ts.step(images, labels) for _ in range(4)])
        if kind == "conv":
            assert rel_err(conv.convolution.params.reshape(32, 16).T, np.asarray(flat.fullconnect.params)) < TOL
    assert np.allclose(steps[0], steps[1], rtol=1e-5)
    assert steps[0][-1] < steps[0][0]


def test_remaining_activations_and_losses(nt):
    """sigmoid / tanh / SiLU (net/activation.py:25-73), BCE on logits and MSE (net/loss.py:53-57, 68-71) vs the vectors of
    the unmodified reference, plus a larger random case vs the oracle (vectorised path, non-multiple-of-4 size)."""
    from net.activation import sigmoid, tanh, SiLU
    from net.loss import binary_cross_entropy_logit_loss, mean_square_loss
    g = load_golden("conv.npz")
    x, dy = g["act_x"], g["act_dy"]
    for name, cls in (("sigmoid", sigmoid), ("tanh", tanh), ("silu", SiLU)):
        a = cls()
        assert rel_err(a.forward(x), g[name + "_y"]) < TOL
        assert rel_err(a.backward(dy), g[name + "_dx"]) < TOL
        assert not hasattr(a, "save_model")                     # checkpoint walkers must keep skipping activations
    loss, partial, s = binary_cross_entropy_logit_loss(x, g["bce_label"])
    assert abs(float(loss) - g["bce_loss"]) < TOL * abs(g["bce_loss"])
    assert rel_err(partial, g["bce_partial"]) < TOL and rel_err(s, g["bce_sigmoid"]) < TOL
    loss, partial = mean_square_loss(x, g["mse_label"])
    assert abs(float(loss) - g["mse_loss"]) < TOL * abs(g["mse_loss"]) and rel_err(partial, g["mse_partial"]) < TOL
    rs = np.random.RandomState(2)
    big, gy, lab = rs.randn(333, 1001) * 3, rs.randn(333, 1001), (rs.rand(333, 1001) > 0.3).astype(np.float64)
    a = SiLU()
    assert rel_err(a.forward(big), R.silu_fwd(big)) < TOL and rel_err(a.backward(gy), R.silu_bwd(gy, big)) < TOL
    loss, partial, s = binary_cross_entropy_logit_loss(big, lab)
    rl, rp, rsig = R.bce_logit_loss(big, lab)
    assert abs(float(loss) - rl) < TOL * abs(rl) and rel_err(partial, rp) < TOL and rel_err(s, rsig) < TOL
