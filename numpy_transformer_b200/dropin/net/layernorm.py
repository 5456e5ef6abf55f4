"""net.layernorm.layer_norm on the B200 backend (mirrors net/layernorm.py:39-164)."""
import numpy as np
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops


class layer_norm(object):
    def __init__(self, normalized_shape, elementwise_affine=True, gamma=[], beta=[], adam=False, float32=False,
                 float16=False):
        if isinstance(normalized_shape, int):
            normalized_shape = [normalized_shape]
        self.normalized_shape = list(normalized_shape)
        self.elementwise_affine = elementwise_affine
        self.host_dtype = np.float32 if float32 else (np.float16 if float16 else np.float64)
        D = int(np.prod(self.normalized_shape))
        g = np.asarray(gamma, dtype=np.float64).reshape(-1) if (elementwise_affine and len(gamma)) else np.ones(D)
        b = np.asarray(beta, dtype=np.float64).reshape(-1) if (elementwise_affine and len(beta)) else np.zeros(D)
        self.gamma = DeviceArray.from_host(g, host_dtype=self.host_dtype)
        self.beta = DeviceArray.from_host(b, host_dtype=self.host_dtype)
        self.gamma_delta = DeviceArray.zeros((D,), host_dtype=self.host_dtype)
        self.beta_delta = DeviceArray.zeros((D,), host_dtype=self.host_dtype)
        self.adam = adam
        self.beta1, self.beta2, self.epsadam = 0.9, 0.999, 10 ** (2 - 10)
        self.moment_g = self.rmsprop_g = self.moment_b = self.rmsprop_b = None
        if adam and elementwise_affine:
            self.moment_g, self.rmsprop_g = DeviceArray.zeros((D,)), DeviceArray.zeros((D,))
            self.moment_b, self.rmsprop_b = DeviceArray.zeros((D,)), DeviceArray.zeros((D,))
        self.t = 1
        self.ep = 1e-5
        self._lead = 0          # the reference expands gamma/beta to the input rank on first forward (:105-108)
        self.inputs = self.mean = self.rstd = None

    def _forward(self, x, post_add=None, split_out=False):
        """split_out: the output feeds a Linear; emit its bf16 hi/lo planes too when that Linear will run BF16x3."""
        self.inputs = x
        self._lead = x.ndim - len(self.normalized_shape)
        y, self.mean, self.rstd = ops.layernorm_fwd(x, self.gamma, self.beta, post_add, split_out=split_out)
        return y

    def _backward(self, dy, addend=None, want_dx=True, split_out=False):
        aff = self.elementwise_affine
        return ops.layernorm_bwd(dy, self.inputs, self.mean, self.rstd, self.gamma,
                                 self.gamma_delta if aff else None, self.beta_delta if aff else None, addend, want_dx,
                                 split_out=split_out)

    def forward(self, inputs):
        return self._forward(as_device(inputs, self.host_dtype))

    def backward(self, delta):
        return self._backward(as_device(delta, self.host_dtype))

    def setzero(self):
        self.gamma_delta.fill_zero()
        self.beta_delta.fill_zero()

    def update(self, lr=1e-10):
        if not self.elementwise_affine:
            return
        if self.adam:
            ops.adam_star(self.gamma, self.gamma_delta, self.moment_g, self.rmsprop_g, lr, self.t)
            ops.adam_star(self.beta, self.beta_delta, self.moment_b, self.rmsprop_b, lr, self.t)
            self.t += 1
        else:
            ops.sgd(self.gamma, self.gamma_delta, lr)
            ops.sgd(self.beta, self.beta_delta, lr)

    def _host_shape(self):
        return (1,) * self._lead + tuple(self.normalized_shape)

    def save_model(self):
        return [np.asarray(self.gamma).reshape(self._host_shape()), np.asarray(self.beta).reshape(self._host_shape())]

    def _check_model(self, models):
        """layernorm.py:162-164 reshapes models[0] / models[1] to the parameter shapes: same errors, nothing overwritten."""
        np.asarray(models[0]).reshape(self.gamma.shape)
        np.asarray(models[1]).reshape(self.beta.shape)

    def restore_model(self, models):
        self._check_model(models)
        self.gamma.set(np.asarray(models[0]).reshape(-1))
        self.beta.set(np.asarray(models[1]).reshape(-1))

    def _slots(self):
        if not self.elementwise_affine:
            return []
        return [(self, "gamma", "gamma_delta", "moment_g", "rmsprop_g"),
                (self, "beta", "beta_delta", "moment_b", "rmsprop_b")]
