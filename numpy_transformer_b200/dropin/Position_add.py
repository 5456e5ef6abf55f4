"""Position_add.Position_learnable — ViT positional add on the B200 backend (mirrors Position_add.py:4-49).
fixed=1 (README config) adds the sin/cos table; otherwise a learned (1,N,D) parameter."""
import numpy as np
from gpt.FixedEmbed import sincos_table
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops


class Position_learnable():
    def __init__(self, n_patch, embed_dim, fixed=False):
        n = n_patch ** 2
        self.n, self.embed_dim = n, embed_dim
        self.param = DeviceArray.from_host(np.random.normal(0, 1, (1, n, embed_dim)))      # drawn even when fixed (:6)
        self.param_delta = DeviceArray.zeros((1, n, embed_dim))
        self.fixed = fixed
        self.posk = DeviceArray.from_host(sincos_table(n, embed_dim)[None])

    def table(self):
        return self.posk if self.fixed else self.param

    def forward(self, inputs):
        x = as_device(inputs)
        out = x.empty_like()
        lib.nt_add_rows(x.ptr, self.table().ptr, out.ptr, x.size // self.embed_dim, self.n, self.embed_dim)
        return out

    def backward(self, delta):
        d = as_device(delta)
        if not self.fixed:
            lib.nt_embedding_bwd(None, d.ptr, None, self.param_delta.ptr, d.shape[0], self.n, self.embed_dim, 1)
        return d

    def update(self, lr):
        if not self.fixed:
            ops.sgd(self.param, self.param_delta, lr)

    def setzero(self):
        if not self.fixed:
            self.param_delta.fill_zero()

    def save_model(self):
        return [] if self.fixed else [np.asarray(self.param)]

    def restore_model(self, models):
        if not self.fixed:
            self.param.set(np.asarray(models[0]).reshape(self.param.shape))

    def _slots(self):
        return [] if self.fixed else [(self, "param", "param_delta", None, None)]
