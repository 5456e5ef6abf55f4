"""net.flatten.flatten_layer (mirrors net/flatten.py:3-15): a view, no parameters, no save_model."""
from numpy_transformer_b200.device import as_device


class flatten_layer():
    def forward(self, inputs):
        x = as_device(inputs)
        self.shape = x.shape
        return x.reshape(x.shape[0], -1)

    def backward(self, delta, lr=''):
        return as_device(delta).reshape(self.shape)

    def update(self, lr=''):
        pass

    def setzero(self):
        pass
