"""gpt.FixedEmbed.Position_Fixed — fixed sin/cos positional lookup, no parameters
(mirrors gpt/FixedEmbed.py:11-40)."""
import numpy as np
from numpy_transformer_b200.device import DeviceArray


def sincos_table(n, D):
    """posk[i,j] = sin(i/10000^(j/D)) (j even) | cos(i/10000^((j-1)/D)) (j odd); FixedEmbed.py:13-19."""
    i = np.arange(n, dtype=np.float64)[:, None]
    j = np.arange(D)[None, :]
    je = np.where(j % 2 == 0, j, j - 1).astype(np.float64)
    ang = i / (10000.0 ** (je / D))
    return np.where(j % 2 == 0, np.sin(ang), np.cos(ang))


class Position_Fixed():
    def __init__(self, context_length, embed_dim):
        self.posk = sincos_table(context_length, embed_dim)

    def forward(self, inputs):
        return self.posk[np.asarray(inputs), :]

    def backward(self, delta):
        return delta

    def update(self, lr):
        return

    def setzero(self):
        return

    def save_model(self):
        return []

    def restore_model(self, models):
        return
