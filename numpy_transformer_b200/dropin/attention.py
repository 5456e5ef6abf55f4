"""attention.attention_layer — the ViT encoder block on the B200 backend (mirrors attention.py:6-114).

  h   = x + MHA(LN(x))                 (no output projection)
  out = h + FC1(GELU(FC0(LN1(h))))     (MLP ratio 2)

The reference's per-sample / per-head Python loops (attention.py:31-45, 63-83) become one fused
attention kernel each way; GELU, bias and both residual adds ride in GEMM / attention epilogues."""
import numpy as np
from net.layernorm import layer_norm
from net.fullconnect import fclayer
from net.activation import Softmax, GELU
from numpy_transformer_b200.device import as_device
from numpy_transformer_b200 import ops
from numpy_transformer_b200 import fused_vit
from numpy_transformer_b200.layer_arena import LayerArenaMixin


class attention_layer(LayerArenaMixin):
    def __init__(self, embed_dim, num_h):
        self.embed_dim = embed_dim
        self.num_h = num_h
        self.len_single = embed_dim // num_h
        self.norm = layer_norm(self.embed_dim)
        self.norm1 = layer_norm(self.embed_dim)
        self.qkvfc = fclayer(self.embed_dim, self.len_single * num_h * 3, True)
        self.fc0 = fclayer(self.embed_dim, self.embed_dim * 2, True)
        self.fc1 = fclayer(self.embed_dim * 2, self.embed_dim, True)
        self.softmax = Softmax()
        self.gelu = GELU()

    def forward(self, inputs, masks=[]):
        x = as_device(inputs)
        self.batch, self.block, _ = x.shape
        kind, dmask = ops.detect_mask(masks, self.block)
        self._mask_kind = kind
        self._fused = False
        if kind == ops.MASK_NONE and fused_vit.supported(self.block, self.embed_dim, self.num_h):
            return fused_vit.forward_stack([self], x)             # whole block in one launch (csrc/vit_block.cu)
        # split_out: producers whose output feeds a Linear also emit its bf16 hi/lo planes (BF16x3 engine, large shapes)
        self.out = self.norm._forward(x, split_out=True)
        self.qkv = self.qkvfc._forward(self.out, split_out=ops.attention_wants_planes(       # (B,N,3D): [q,k,v][head][dh]
            self.block, self.embed_dim // self.num_h, self._mask_kind))
        input1, self.P, _ = ops.attention_fwd(self.qkv, self.num_h, dmask, kind, residual=x)
        self.out1 = self.norm1._forward(input1, split_out=True)
        self.out2, self.pre2 = self.fc0._forward(self.out1, ops.ACT_GELU, want_pre=True, split_out=True)
        return self.fc1._forward(self.out2, residual=input1)          # out6 + input1

    def backward(self, delta):
        delta = as_device(delta)
        if getattr(self, "_fused", False):
            return fused_vit.backward_stack([self], delta)
        d = self.fc1._backward(delta, self.out2, ops.ACT_GELU, self.pre2, split_out=True)   # fc1.backward + gelu.backward
        d = self.fc0._backward(d, self.out1)
        delta0 = self.norm1._backward(d, addend=delta)                 # delta0 += delta (attention.py:62)
        self.delta_qkv = ops.attention_bwd(delta0, self.qkv, self.P, self.num_h, self._mask_kind, split_out=True)
        d = self.qkvfc._backward(self.delta_qkv, self.out)
        return self.norm._backward(d, addend=delta0, split_out=True)   # input_delta += delta0 (attention.py:88)

    @property
    def atg__(self):
        """[batch][head] -> (N,N) softmax matrices, as the reference keeps them (attention.py:30,41)."""
        P = np.asarray(self.P)
        return [[P[n, i] for i in range(self.num_h)] for n in range(self.batch)]

    def _children(self):
        return [self.norm, self.qkvfc, self.norm1, self.fc0, self.fc1]

    def update(self, lr):
        if self._arena_update(lr):                        # one kernel for the block's parameter tensors
            return
        for c in (self.norm, self.norm1, self.qkvfc, self.fc0, self.fc1):
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
