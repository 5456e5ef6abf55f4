"""Autoregressive generation with a KV cache.

The reference's samplers (gpt/gpt_predict_poetrythree.py:108-121; the every-50-steps sample inside
gpt/gpt_train_potry3000.py:300-342) re-run the whole window through every layer for each new token.  While the window
is still growing nothing about the earlier tokens changes (causal mask; learned positions are fixed per index), so
`KVGenerator` keeps each block's keys / values on the device and a step only pushes ONE token through the stack:
LayerNorm -> FC0 -> ReLU -> FC1 (+res) -> LayerNorm -> qkv -> append k,v -> attention over the cache -> fc_out (+res).
Once the window is full the reference slides it (every position index changes), so `generate` falls back to the
full-window recomputation from there on — the same arithmetic the reference performs.

Layers are the drop-in objects of a GPT layer list (Position_Embedding, attdecoderblock_layer ..., layer_norm,
classify_layer(relu=False)); parameters are shared with training, nothing is copied."""
import numpy as np

from ._lib import lib
from .device import DeviceArray
from . import ops


class KVGenerator:
    def __init__(self, layers):
        self.layers = layers
        self.embed = layers[0]
        self.blocks = [l for l in layers if type(l).__name__ == "attdecoderblock_layer"]
        self.final_norm, self.head = layers[-2], layers[-1]
        assert type(self.embed).__name__ == "Position_Embedding" and self.blocks, "KVGenerator needs a GPT layer list"
        assert not getattr(self.head, "reluact", False) or True
        self.context = self.embed.context_length
        self.D = self.embed.embed_dim
        self.len = 0
        self.k = self.v = None

    # ------------------------------------------------------------------ full window (the reference's way)
    def full_logits(self, tokens):
        """Logits of the LAST position from a forward over the whole window (gpt_predict_poetrythree.py:109-114)."""
        x = DeviceArray.from_host(np.asarray(tokens).astype(np.int64))
        for l in self.layers:
            x = l.forward(x, "causal") if l in self.blocks else l.forward(x)
        return np.asarray(x)[:, -1, :]

    # ------------------------------------------------------------------ cached path
    def prefill(self, tokens):
        """Run the prompt [B, T0] through the stack once and fill the caches; returns the last position's logits."""
        tokens = np.asarray(tokens).astype(np.int64)
        B, T0 = tokens.shape
        assert T0 <= self.context
        self.k = [DeviceArray.zeros((B, self.context, self.D)) for _ in self.blocks]
        self.v = [DeviceArray.zeros((B, self.context, self.D)) for _ in self.blocks]
        x = DeviceArray.from_host(tokens)
        bi = 0
        for l in self.layers:
            if l in self.blocks:
                x = l.forward(x, "causal")
                lib.nt_kv_append(l.qkv.ptr, self.k[bi].ptr, self.v[bi].ptr, B, T0, self.context, 0, self.D)
                bi += 1
            else:
                x = l.forward(x)
        self.len = T0
        return np.asarray(x)[:, -1, :]

    def step(self, token):
        """Push one new token per sample ([B] ints) through the stack using the caches; returns its logits [B, V]."""
        assert self.k is not None and self.len < self.context, "prefill first / window is full"
        tok = np.asarray(token).astype(np.int64).reshape(-1, 1)
        B, D, pos = tok.shape[0], self.D, self.len
        idx = DeviceArray.from_host(tok)
        x = DeviceArray((B, 1, D), host_dtype=self.embed.text_embedding.host_dtype)
        lib.nt_embedding_fwd(idx.ptr, self.embed.text_embedding.params.ptr, self.embed.pos_embedding.params.ptr + 4 * pos * D,
                             x.ptr, B, 1, D)                       # token row + position row `pos` (PatchEmbed.py:132-138)
        for bi, blk in enumerate(self.blocks):
            u = blk.norm._forward(x)
            hdn = blk.fc0._forward(u, ops.ACT_RELU)
            h = blk.fc1._forward(hdn, residual=x)
            u1 = blk.norm1._forward(h)
            qkv = blk.qkvfc._forward(u1)                           # [B,1,3D]
            lib.nt_kv_append(qkv.ptr, self.k[bi].ptr, self.v[bi].ptr, B, 1, self.context, pos, D)
            att = DeviceArray((B, 1, D), host_dtype=x.host_dtype)
            lib.nt_attention_decode(qkv.ptr, self.k[bi].ptr, self.v[bi].ptr, att.ptr, B, self.context, pos + 1, blk.num_h,
                                    D // blk.num_h)
            x = blk.fc_out._forward(att, residual=h)
        z = self.final_norm._forward(x)
        logits = self.head.forward(z)
        self.len = pos + 1
        return np.asarray(logits).reshape(B, -1)

    def generate(self, prompt, n_new, sample=None):
        """Greedy (sample=None) or sampled (sample(probs [B,V]) -> [B] ints) continuation of prompt [B, T0]."""
        out = [list(map(int, row)) for row in np.asarray(prompt)]
        logits = self.prefill(np.asarray(prompt)[:, -self.context:])
        for _ in range(n_new):
            if sample is None:
                nxt = np.argmax(logits, axis=-1)
            else:
                e = np.exp(logits - logits.max(-1, keepdims=True))
                nxt = np.asarray(sample(e / e.sum(-1, keepdims=True)))
            for row, t in zip(out, nxt):
                row.append(int(t))
            if self.len < self.context:
                logits = self.step(nxt)
            else:                                                  # window full: it slides, every position index changes
                logits = self.full_logits(np.array([row[-self.context:] for row in out]))
        return np.array(out)
