"""PatchEmbed.{PatchEmbed_flatten, Position_Embedding, PatchEmbed_convolution} on the B200 backend
(mirrors PatchEmbed.py:10-71, 124-159).  The three Python patch loops become one gather kernel; token +
position embedding is one gather kernel forward and two scatter/reduce kernels backward."""
import numpy as np
from gpt.FixedEmbed import Position_Fixed  # noqa: F401  (the reference imports it here too)
from net.fullconnect import fclayer
from net.Convolution import convolution_layer  # noqa: F401
from net.layernorm import layer_norm
from net.Embedding import Embedding_layer
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray, as_device


class PatchEmbed_flatten(object):
    def __init__(self, embed_dim, images_shape, n_patch, patchnorm=True):
        self.embed_dim = embed_dim
        n, c, h, w = images_shape
        self.h_length, self.w_length = h // n_patch, w // n_patch
        self.n_patch, self.patchnorm = n_patch, patchnorm
        self.fullconnect = fclayer(self.h_length * self.w_length * c, self.embed_dim, True)
        if patchnorm:
            self.norm = layer_norm(self.embed_dim)

    def _patches(self, images):
        x = as_device(images)
        n, c, h, w = x.shape
        out = DeviceArray((n, self.n_patch ** 2, self.h_length * self.w_length * c))
        lib.nt_patchify(x.ptr, out.ptr, n, c, h, w, self.n_patch)
        return out

    def forward(self, images, post_add=None):
        self.inputs = self._patches(images)
        out = self.fullconnect._forward(self.inputs)
        if self.patchnorm:
            out = self.norm._forward(out, post_add)
        return out

    def backward(self, delta, want_dx=True):
        delta = as_device(delta)
        if self.patchnorm:
            delta = self.norm._backward(delta)
        return self.fullconnect._backward(delta, self.inputs, want_dx=want_dx)   # (B,49,16) patch-space grad

    def update(self, lr):
        if self.patchnorm:
            self.norm.update(lr)
        self.fullconnect.update(lr)

    def setzero(self):
        if self.patchnorm:
            self.norm.setzero()
        self.fullconnect.setzero()

    def save_model(self):
        model = []
        if self.patchnorm:
            model.append(self.norm.save_model())
        model.append(self.fullconnect.save_model())
        return model

    def restore_model(self, models):
        if self.patchnorm:
            self.norm._check_model(models[0])
        self.fullconnect._check_model(models[-1])
        if self.patchnorm:
            self.norm.restore_model(models[0])
        self.fullconnect.restore_model(models[-1])

    def _slots(self):
        return (self.norm._slots() if self.patchnorm else []) + self.fullconnect._slots()


class PatchEmbed_convolution(object):
    """PatchEmbed.py:73-122: convolution with kernel = stride = patch size, then (n, patches, embed_dim).  As in the
    reference, forward runs the LayerNorm but hands back the tensor BEFORE it (:90-93 returns `out`), while backward
    pushes the incoming gradient through the norm's backward; both quirks are kept (tests/golden/conv.npz)."""

    def __init__(self, embed_dim, images_shape, n_patch, patchnorm=True):
        self.embed_dim = embed_dim
        n, c, h, w = images_shape
        self.batch = n
        self.h_length, self.w_length = h // n_patch, w // n_patch
        self.n_patch, self.patchnorm = n_patch, patchnorm
        self.convolution = convolution_layer(c, embed_dim, kernel_size=self.w_length, stride=self.w_length)
        if patchnorm:
            self.norm = layer_norm(self.embed_dim)

    def forward(self, images):
        out = self.convolution._forward_nhwc(images)           # already (n, ph*pw, embed_dim): no NCHW round trip
        if self.patchnorm:
            self.norm._forward(out)
        return out

    def backward(self, delta, want_dx=True):
        delta = as_device(delta)
        if self.patchnorm:
            delta = self.norm._backward(delta)
        return self.convolution._backward_nhwc(delta, want_dx=want_dx)

    def update(self, lr):
        if self.patchnorm:
            self.norm.update(lr)
        self.convolution.update(lr)

    def setzero(self):
        if self.patchnorm:
            self.norm.setzero()
        self.convolution.setzero()

    def save_model(self):
        model = []
        if self.patchnorm:
            model.append(self.norm.save_model())
        model.append(self.convolution.save_model())
        return model

    def restore_model(self, models):
        if self.patchnorm:
            self.norm.restore_model(models[0])
        self.convolution.restore_model(models[-1])

    def _slots(self):
        return (self.norm._slots() if self.patchnorm else []) + self.convolution._slots()


class Position_Embedding(Embedding_layer):
    def __init__(self, context_length, vocab_size, embed_dim, adam=False, float32=False, float16=False):
        self.context_length = context_length
        self.embed_dim = embed_dim
        kw = dict(adam=adam, float32=float32, float16=float16)
        self.text_embedding = Embedding_layer(vocab_size, embedding_dim=embed_dim, **kw)
        self.pos_embedding = Embedding_layer(context_length, embedding_dim=embed_dim, **kw)
        self.adam = adam

    def forward(self, inputs):
        idx = self.text_embedding._idx(inputs)
        n, T = idx.shape
        self.idx = idx
        out = DeviceArray((n, T, self.embed_dim), host_dtype=self.text_embedding.host_dtype)
        lib.nt_embedding_fwd(idx.ptr, self.text_embedding.params.ptr, self.pos_embedding.params.ptr, out.ptr,
                             n, T, self.embed_dim, self.text_embedding.num_embeddings)
        return out

    def backward(self, delta):
        d = as_device(delta)
        n, T = self.idx.shape
        lib.nt_embedding_bwd(self.idx.ptr, d.ptr, self.text_embedding.delta.ptr, self.pos_embedding.delta.ptr,
                             n, T, self.embed_dim, self.text_embedding.num_embeddings)
        return self.text_embedding.delta

    def update(self, lr):
        self.text_embedding.update(lr)
        self.pos_embedding.update(lr)

    def setzero(self):
        self.text_embedding.setzero()
        self.pos_embedding.setzero()

    def save_model(self):
        return [self.text_embedding.save_model(), self.pos_embedding.save_model()]

    def restore_model(self, models):
        self.text_embedding._check_model(models[0])
        self.pos_embedding._check_model(models[1])
        self.text_embedding.restore_model(models[0])
        self.pos_embedding.restore_model(models[1])

    def _slots(self):
        return self.text_embedding._slots() + self.pos_embedding._slots()
