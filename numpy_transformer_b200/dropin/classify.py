"""classify.classify_layer — FC0 -> (ReLU) -> FC1 head on the B200 backend (mirrors classify.py:6-55)."""
import numpy as np
from net.layernorm import layer_norm
from net.fullconnect import fclayer
from net.activation import ReLU
from numpy_transformer_b200.device import DeviceArray, as_device
from numpy_transformer_b200 import ops


class classify_layer():
    def __init__(self, embed_dim, batch, n_patch, num_classes, cls_token=True, adam=False, relu=True, float32=False,
                 float16=False):
        self.batch, self.embed_dim, self.n_patch, self.cls_token = batch, embed_dim, n_patch, cls_token
        kw = dict(adam=adam, float32=float32, float16=float16)
        if cls_token:
            self.fc0 = fclayer(self.embed_dim, self.embed_dim, True, **kw)
        else:
            self.fc0 = fclayer(self.embed_dim * int(n_patch ** 2), self.embed_dim, True, **kw)
        self.relu = ReLU()
        self.fc1 = fclayer(self.embed_dim, num_classes, True, **kw)
        self.reluact = relu

    def forward(self, inputs):
        x = as_device(inputs, self.fc0.host_dtype)
        self.inputs = x
        if self.reluact:
            self.out1, self.pre1 = self.fc0._forward(x, ops.ACT_RELU, want_pre=True)
        else:
            self.out1, self.pre1 = self.fc0._forward(x), None
        return self.fc1._forward(self.out1)

    def backward(self, delta):
        delta = as_device(delta, self.fc0.host_dtype)
        if self.reluact:
            d = self.fc1._backward(delta, self.out1, ops.ACT_RELU, self.pre1)
        else:
            d = self.fc1._backward(delta, self.out1)
        d = self.fc0._backward(d, self.inputs)
        if self.cls_token:
            # classify.py:35-38: gradient lands on token 0 only (non-README switch; host scatter)
            zeros = np.zeros((self.batch, int(self.n_patch ** 2), self.embed_dim))
            zeros[:, 0] = np.asarray(d)
            return DeviceArray.from_host(zeros)
        return d

    def update(self, lr):
        self.fc1.update(lr)
        self.fc0.update(lr)

    def setzero(self):
        self.fc0.setzero()
        self.fc1.setzero()

    def save_model(self):
        return [self.fc0.save_model(), self.fc1.save_model()]

    def _check_model(self, models):
        self.fc0._check_model(models[0])
        self.fc1._check_model(models[1])

    def restore_model(self, models):
        self._check_model(models)
        self.fc0.restore_model(models[0])
        self.fc1.restore_model(models[1])

    def _slots(self):
        return self.fc0._slots() + self.fc1._slots()
