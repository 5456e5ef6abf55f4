"""net.Convolution.convolution_layer on the B200 backend (mirrors net/Convolution.py:37-248).

The reference computes the forward as im2col + matmul (:120-131) and the backward as two more im2col + matmul passes over
a zero-stuffed / rotated formulation (:163-225).  Here the layer is an `nt_im2col` gather followed by the SAME Linear
engine every other layer uses (an internal fclayer over the (C*KH*KW, OC) view of the kernel tensor: tcgen05 GEMMs with
fused bias, split-K weight gradient with fused bias gradient), and `nt_col2im` — the adjoint gather — for the input
gradient.  `params` / `params_delta` keep the reference's (OC, C, KH, KW) layout at the API surface and in pickles."""
import numpy as np
from net.fullconnect import fclayer
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device


def _pair(v):
    return [int(v), int(v)] if isinstance(v, (int, np.integer)) else [int(v[0]), int(v[1])]


class convolution_layer(object):
    def __init__(self, in_channel, out_channel, kernel_size, stride=[1, 1], padding=[0, 0], bias=True, params=[],
                 bias_params=[]):
        self.in_channel, self.out_channel = in_channel, out_channel
        self.kernel_size = _pair(kernel_size)
        self.bias = bias
        kh, kw = self.kernel_size
        K = in_channel * kh * kw
        # same draws, in the same order, as net/Convolution.py:45-55 (the bias is drawn even if unused)
        r = np.sqrt(1 / K)
        w = np.asarray(params) if len(params) else np.random.uniform(-r, r, (out_channel, in_channel, kh, kw))
        b = np.asarray(bias_params) if (bias and len(bias_params)) else np.random.uniform(-r, r, (out_channel))
        self._fc = fclayer(K, out_channel, bias, params=np.ascontiguousarray(w.reshape(out_channel, K).T), bias_params=b)
        self.stride = _pair(stride)
        self.padding = padding if isinstance(padding, str) else _pair(padding)
        self.outshape = self.ishape = None
        self._col = None

    # ---- the reference's attribute surface: (OC, C, KH, KW) host views of the Linear-shaped device tensors
    def _kernel_view(self, dev):
        kh, kw = self.kernel_size
        return np.ascontiguousarray(np.asarray(dev).T).reshape(self.out_channel, self.in_channel, kh, kw)

    def _kernel_store(self, dev, value):
        dev.set(np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(self.out_channel, -1).T))
        self._fc._wsplit = None

    params = property(lambda self: self._kernel_view(self._fc.params), lambda self, v: self._kernel_store(self._fc.params, v))
    params_delta = property(lambda self: self._kernel_view(self._fc.params_delta),
                            lambda self, v: self._kernel_store(self._fc.params_delta, v))
    bias_params = property(lambda self: self._fc.bias_params, lambda self, v: self._fc.bias_params.set(np.asarray(v)))
    bias_delta = property(lambda self: self._fc.bias_delta, lambda self, v: self._fc.bias_delta.set(np.asarray(v)))

    # ---- forward / backward in the GEMM's own layout: rows = (n, oh, ow), columns = channels
    def _geometry(self, shape):
        if isinstance(self.padding, str):               # net/Convolution.py:107-111
            if self.padding == 'same':
                self.padding = [self.kernel_size[0] // 2, self.kernel_size[1] // 2]
            elif self.padding == 'valid':
                self.padding = [0, 0]
            else:
                raise ValueError("convolution_layer: padding %r" % (self.padding,))
        n, c, h, w = shape
        assert c == self.in_channel, "convolution_layer: %d input channels, expected %d" % (c, self.in_channel)
        self.ishape = tuple(shape)
        self.oh = (h + 2 * self.padding[0] - self.kernel_size[0]) // self.stride[0] + 1
        self.ow = (w + 2 * self.padding[1] - self.kernel_size[1]) // self.stride[1] + 1
        self.outshape = (n, self.out_channel, self.oh, self.ow)

    def _conv_args(self):
        n, c, h, w = self.ishape
        return (n, c, h, w, self.kernel_size[0], self.kernel_size[1], self.stride[0], self.stride[1], self.padding[0],
                self.padding[1])

    def _forward_nhwc(self, inputs):
        """-> [N, OH*OW, OC] (the layout PatchEmbed_convolution wants after its transpose, PatchEmbed.py:91-92)."""
        x = as_device(inputs)
        self.inputs = x
        self._geometry(x.shape)
        n = x.shape[0]
        K = self.in_channel * self.kernel_size[0] * self.kernel_size[1]
        col = DeviceArray((n * self.oh * self.ow, K), host_dtype=x.host_dtype)
        lib.nt_im2col(x.ptr, col.ptr, *self._conv_args())
        self._col = col
        return self._fc._forward(col).reshape(n, self.oh * self.ow, self.out_channel)

    def _backward_nhwc(self, delta, want_dx=True):
        """delta [N, OH*OW, OC] -> dx [N, C, H, W]; accumulates params_delta / bias_delta (Convolution.py:165-166, 185-192)."""
        d2 = delta.reshape(-1, self.out_channel)
        dcol = self._fc._backward(d2, self._col, want_dx=want_dx)
        if not want_dx:
            return None
        dx = DeviceArray(self.ishape, host_dtype=delta.host_dtype)
        lib.nt_col2im(dcol.ptr, dx.ptr, *self._conv_args())
        return dx

    # ---- reference protocol (NCHW in, NCHW out)
    def forward(self, inputs):
        y = self._forward_nhwc(inputs)
        n = y.shape[0]
        out = DeviceArray(self.outshape, host_dtype=y.host_dtype)
        lib.nt_transpose_last2(y.ptr, out.ptr, n, self.oh * self.ow, self.out_channel)
        return out

    def backward(self, delta):
        d = as_device(delta)
        n = self.outshape[0]
        d2 = DeviceArray((n, self.oh * self.ow, self.out_channel), host_dtype=d.host_dtype)
        lib.nt_transpose_last2(d.ptr, d2.ptr, n, self.out_channel, self.oh * self.ow)
        return self._backward_nhwc(d2)

    def getdelta(self):
        return [] if self.bias else [self.params_delta]

    def setzero(self):
        self._fc.setzero()
        if not self.bias:
            self._fc.bias_delta.fill_zero()

    def update(self, lr=1e-10):
        self._fc.update(lr)                              # SGD, net/Convolution.py:236-239

    def save_model(self):
        return [self.params, np.asarray(self.bias_params)] if self.bias else [self.params]

    def restore_model(self, models):
        self.params = np.asarray(models[0])
        if self.bias:
            self._fc.bias_params.set(np.asarray(models[1]).reshape(-1))

    def _slots(self):
        return self._fc._slots()
