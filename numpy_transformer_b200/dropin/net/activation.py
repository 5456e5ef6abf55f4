"""net.activation.{ReLU, sigmoid, tanh, SiLU, Softmax, GELU} on the B200 backend (mirrors net/activation.py:9-154).
As in the reference these classes have no save_model/restore_model, so checkpoint walkers skip them."""
import numpy as np
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device


class ReLU(object):
    def forward(self, inputs):
        host = inputs if isinstance(inputs, np.ndarray) else None
        x = as_device(inputs)
        self.inputs = x.copy()                      # deepcopy of the pre-activation (activation.py:12)
        lib.nt_relu_fwd(x.ptr, x.ptr, x.size)       # in place, returns the same object (activation.py:13-14)
        if host is not None:
            host[host < 0] = 0
        return x

    def backward(self, delta):
        d = as_device(delta)
        out = d.empty_like()
        lib.nt_relu_bwd(d.ptr, self.inputs.ptr, out.ptr, d.size)
        return out

    def setzero(self):
        pass

    def update(self, lr=1e-10):
        pass


class _Elementwise(object):
    """y = f(x); backward needs either the output (sigmoid, tanh) or the input (SiLU)."""
    _fwd = _bwd = None
    _keep_output = True

    def forward(self, inputs):
        x = as_device(inputs)
        self.inputs = x
        self.y = x.empty_like()
        getattr(lib, self._fwd)(x.ptr, self.y.ptr, x.size)
        return self.y

    def backward(self, delta):
        d = as_device(delta)
        out = d.empty_like()
        getattr(lib, self._bwd)(d.ptr, (self.y if self._keep_output else self.inputs).ptr, out.ptr, d.size)
        return out

    def setzero(self):
        pass

    def update(self, lr=1e-10):
        pass


class sigmoid(_Elementwise):          # net/activation.py:25-39
    _fwd, _bwd = "nt_sigmoid_fwd", "nt_sigmoid_bwd"


class tanh(_Elementwise):             # net/activation.py:41-56
    _fwd, _bwd = "nt_tanh_fwd", "nt_tanh_bwd"


class SiLU(_Elementwise):             # net/activation.py:58-73
    _fwd, _bwd, _keep_output = "nt_silu_fwd", "nt_silu_bwd", False


class GELU(object):
    def forward(self, inputs):
        x = as_device(inputs)
        self.inputs = x
        out = x.empty_like()
        lib.nt_gelu_fwd(x.ptr, out.ptr, x.size)
        return out

    def backward(self, delta):
        d = as_device(delta)
        out = d.empty_like()
        lib.nt_gelu_bwd(d.ptr, self.inputs.ptr, out.ptr, d.size)
        return out

    def setzero(self):
        pass

    def update(self, lr=1e-10):
        pass


class Softmax(object):
    def forward(self, inputs, axis):
        x = as_device(inputs)
        if axis not in (-1, x.ndim - 1):
            raise ValueError("Softmax: only the last axis is supported on the B200 backend")
        self.axis = axis
        cols = x.shape[-1]
        self.out = x.empty_like()
        lib.nt_softmax_fwd(x.ptr, self.out.ptr, x.size // cols, cols)
        return self.out

    def backward(self, delta, out):
        # reference: 2-D `out`, axis in (1,-1) or exit(-1) (activation.py:95-96); here a RuntimeError
        o = as_device(out)
        d = as_device(delta)
        if o.ndim != 2 or self.axis not in (1, -1):
            raise RuntimeError("Softmax.backward: expects a 2-D softmax output and axis=-1")
        res = o.empty_like()
        res.host_dtype = np.dtype(np.float64)       # activation.py:98
        lib.nt_softmax_bwd(d.ptr, o.ptr, res.ptr, o.shape[0], o.shape[1])
        return res

    def setzero(self):
        pass

    def update(self, lr=1e-10):
        pass
