"""gpt.gpt_linear.gpt_linear_layer — FC0(D->expand*D) -> LN -> FC1 head (mirrors gpt/gpt_linear.py:6-44)."""
from net.layernorm import layer_norm
from net.fullconnect import fclayer
from numpy_transformer_b200.device import as_device


class gpt_linear_layer():
    def __init__(self, embed_dim, num_classes, expand=3):
        self.embed_dim = embed_dim
        self.fc0 = fclayer(self.embed_dim, self.embed_dim * expand, True)
        self.norm = layer_norm(self.embed_dim * expand)
        self.fc1 = fclayer(self.embed_dim * expand, num_classes, True)

    def forward(self, inputs):
        self.inputs = as_device(inputs)
        out0 = self.fc0._forward(self.inputs)
        self.out1 = self.norm._forward(out0)
        return self.fc1._forward(self.out1)

    def backward(self, delta):
        d = self.fc1._backward(as_device(delta), self.out1)
        d = self.norm._backward(d)
        return self.fc0._backward(d, self.inputs)

    def update(self, lr):
        self.fc1.update(lr)
        self.fc0.update(lr)
        self.norm.update(lr)

    def setzero(self):
        self.fc0.setzero()
        self.fc1.setzero()
        self.norm.setzero()

    def save_model(self):
        return [self.fc0.save_model(), self.fc1.save_model(), self.norm.save_model()]

    def restore_model(self, models):
        for c, i in ((self.fc0, 0), (self.fc1, 1), (self.norm, 2)):
            c._check_model(models[i])
        self.fc0.restore_model(models[0])
        self.fc1.restore_model(models[1])
        self.norm.restore_model(models[2])

    def _slots(self):
        return self.fc0._slots() + self.fc1._slots() + self.norm._slots()
