"""net.Embedding.Embedding_layer on the B200 backend (mirrors net/Embedding.py:25-96)."""
import numpy as np
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops


class Embedding_layer(object):
    def __init__(self, num_embeddings, embedding_dim, params=[], adam=False, float32=False, float16=False):
        self.num_embeddings, self.embedding_dim = num_embeddings, embedding_dim
        self.host_dtype = np.float32 if float32 else (np.float16 if float16 else np.float64)
        if len(params):
            w = np.asarray(params)
        else:
            w = np.random.normal(0, 1, (num_embeddings, embedding_dim))      # Embedding.py:32
        self.params = DeviceArray.from_host(w, host_dtype=self.host_dtype)
        self.delta = DeviceArray.zeros((num_embeddings, embedding_dim), host_dtype=np.float64)   # Embedding.py:33
        self.adam = adam
        self.beta1, self.beta2, self.epsadam = 0.9, 0.999, 10 ** (2 - 10)
        self.moment_p = DeviceArray.zeros((num_embeddings, embedding_dim)) if adam else None
        self.rmsprop_p = DeviceArray.zeros((num_embeddings, embedding_dim)) if adam else None
        self.t = 1
        self.flatten = None

    def _idx(self, inputs):
        if isinstance(inputs, DeviceArray):
            assert inputs.dtype == np.int32
            return inputs
        return DeviceArray.from_host(np.asarray(inputs).astype(np.int64))

    def forward(self, inputs):
        idx = self._idx(inputs)
        self.flatten = idx.reshape(-1)
        out = DeviceArray(idx.shape + (self.embedding_dim,), host_dtype=self.host_dtype)
        n = idx.size
        lib.nt_embedding_fwd(idx.ptr, self.params.ptr, None, out.ptr, 1, n, self.embedding_dim, self.num_embeddings)
        return out

    def backward(self, delta, flatten=[]):
        idx = self._idx(flatten).reshape(-1) if len(flatten) else self.flatten
        d = as_device(delta)
        lib.nt_embedding_bwd(idx.ptr, d.ptr, self.delta.ptr, None, 1, idx.size, self.embedding_dim, self.num_embeddings)
        return self.delta

    def setzero(self):
        self.delta.fill_zero()

    def getdelta(self):
        return self.delta

    def normdelta(self, maxnorm, l2norm):
        lib.nt_scale(self.delta.ptr, float(maxnorm) / float(l2norm), self.delta.size)

    def update(self, lr=1e-10):
        if self.adam:
            ops.adam_star(self.params, self.delta, self.moment_p, self.rmsprop_p, lr, self.t)
            self.t += 1
        else:
            ops.sgd(self.params, self.delta, lr)

    def save_model(self):
        return [np.asarray(self.params).astype(np.float32)]          # Embedding.py:93

    def _check_model(self, models):
        np.asarray(models[0]).reshape(self.params.shape)

    def restore_model(self, models):
        self._check_model(models)
        self.params.set(np.asarray(models[0]).reshape(self.params.shape))

    def _slots(self):
        return [(self, "params", "delta", "moment_p", "rmsprop_p")]
