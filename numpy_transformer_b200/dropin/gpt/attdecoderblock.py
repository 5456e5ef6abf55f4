"""gpt.attdecoderblock.attdecoderblock_layer — the GPT decoder block on the B200 backend
(mirrors gpt/attdecoderblock.py:13-148).  MLP first:

  h   = x + FC1(ReLU(FC0(LN(x))))              (ratio 4)
  out = h + FCout(MHA(LN1(h), mask))
"""
import numpy as np
from net.layernorm import layer_norm
from net.fullconnect import fclayer
from net.activation import Softmax, GELU, ReLU
from numpy_transformer_b200.device import as_device
from numpy_transformer_b200 import ops
from numpy_transformer_b200.layer_arena import LayerArenaMixin


class attdecoderblock_layer(LayerArenaMixin):
    def __init__(self, embed_dim, num_h, adam=False, float32=False, float16=False, return_attention=False):
        self.embed_dim = embed_dim
        self.num_h = num_h
        self.len_single = embed_dim // num_h
        kw = dict(adam=adam, float32=float32, float16=float16)
        self.norm = layer_norm(self.embed_dim, **kw)
        self.norm1 = layer_norm(self.embed_dim, **kw)
        self.qkvfc = fclayer(self.embed_dim, self.len_single * num_h * 3, True, **kw)
        self.fc0 = fclayer(self.embed_dim, self.embed_dim * (2 * 2), True, **kw)
        self.fc1 = fclayer(self.embed_dim * (2 * 2), self.embed_dim, True, **kw)
        self.fc_out = fclayer(self.embed_dim, self.embed_dim, True, **kw)
        self.softmax = Softmax()
        self.relu = ReLU()
        self.adam = adam
        self.return_attention = return_attention

    def forward(self, inputs, masks=[]):
        x = as_device(inputs, self.norm.host_dtype)
        self.masks = masks
        self.batch, self.block, _ = x.shape
        kind, dmask = ops.detect_mask(masks, self.block)
        self._mask_kind = kind
        # split_out: producers whose output feeds a Linear also emit its bf16 hi/lo planes (BF16x3 engine, large shapes)
        self.out0 = self.norm._forward(x, split_out=True)
        self.out2, self.pre2 = self.fc0._forward(self.out0, ops.ACT_RELU, want_pre=True, split_out=True)
        self.out6 = self.fc1._forward(self.out2, residual=x)
        self.outnorm = self.norm1._forward(self.out6, split_out=True)
        # qkv also as bf16 planes when the tcgen05 attention forward will take its operands by TMA
        self.qkv = self.qkvfc._forward(self.outnorm, split_out=ops.attention_wants_planes(
            self.block, self.embed_dim // self.num_h, kind, self.return_attention))
        self.rkk, self.P, self.S = ops.attention_fwd(self.qkv, self.num_h, dmask, kind,
                                                     want_scores=self.return_attention, split_out=True)
        input1 = self.fc_out._forward(self.rkk, residual=self.out6)
        if self.return_attention:
            return self.atg__, self.att_col
        return input1

    def backward(self, delta):
        delta = as_device(delta, self.norm.host_dtype)
        # the gradient of the attention output also as bf16 planes when the tcgen05 attention backward will take it by TMA
        delta0 = self.fc_out._backward(delta, self.rkk, split_out=getattr(self.P, "_lse", None) is not None)
        self.delta_qkv = ops.attention_bwd(delta0, self.qkv, self.P, self.num_h, self._mask_kind, split_out=True,
                                           att_out=self.rkk)
        d = self.qkvfc._backward(self.delta_qkv, self.outnorm)
        qkvdelta = self.norm1._backward(d, addend=delta, split_out=True)               # + delta (attdecoderblock.py:105)
        d = self.fc1._backward(qkvdelta, self.out2, ops.ACT_RELU, self.pre2, split_out=True)   # fc1.backward + relu.backward
        d = self.fc0._backward(d, self.out0)
        return self.norm._backward(d, addend=qkvdelta, split_out=True)                 # += qkvdelta (attdecoderblock.py:110)

    @property
    def atg__(self):
        P = np.asarray(self.P)
        return [[P[n, i] for i in range(self.num_h)] for n in range(self.batch)]

    @property
    def att_col(self):
        if self.S is None:
            return [[[] for _ in range(self.num_h)] for _ in range(self.batch)]
        S = np.asarray(self.S)
        return [[S[n, i] for i in range(self.num_h)] for n in range(self.batch)]

    def _children(self):
        return [self.norm, self.qkvfc, self.norm1, self.fc0, self.fc1, self.fc_out]

    def update(self, lr):
        if self._arena_update(lr):                        # one kernel for the block's parameter tensors
            return
        for c in (self.norm, self.norm1, self.qkvfc, self.fc0, self.fc1, self.fc_out):
            c.update(lr)

    def setzero(self):
        if self._arena_setzero():                         # one memset
            return
        for c in self._children():
            c.setzero()

    def save_model(self):
        return [c.save_model() for c in self._children()]

    def _check_model(self, models):
        for i, c in enumerate(self._children()):
            c._check_model(models[i])

    def restore_model(self, models):
        """The reference indexes models[0..] child by child and raises on a tree that is not a block's (a LayerNorm's
        [gamma, beta], say); gpt_train_potry3000.py:219-229 relies on that exception to seed a newly added block from
        the previous one.  Validate the whole tree first so a failed restore leaves the block untouched."""
        self._check_model(models)
        for i, c in enumerate(self._children()):
            c.restore_model(models[i])

    def _slots(self):
        return [s for c in self._children() for s in c._slots()]


def create_masks_future(inputs):
    """gpt/attdecoderblock.py:141-148 (the -1e6 variant)."""
    n, sequence_length = inputs.shape
    input_mask = np.tril(np.ones((sequence_length, sequence_length)))
    input_mask[input_mask == 0] = -1e6
    input_mask[input_mask == 1] = 0
    return input_mask
