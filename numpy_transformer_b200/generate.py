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
import ctypes
import os

import numpy as np

from ._lib import lib
from .device import DeviceArray, keep_allocations, rehome_generation
from . import ops


DECODE_PROGRAM_DEFAULT = "0"        # NTB200_FUSE_DECODE / NTB200_FUSE: "1" = the decode step as one phase program


class KVGenerator:
    """use_graph: the decode step reads its position from a device scalar, so after one eager step (sizes the pool,
    loads modules) it is captured once into a CUDA graph and every later token is one graph launch; greedy
    continuations (`generate(..., sample=None)`) also pick the next token on the device and replay without any host
    round trip until the window is full."""

    def __init__(self, layers, use_graph=True):
        self.layers = layers
        self.embed = layers[0]
        self.blocks = [l for l in layers if type(l).__name__ == "attdecoderblock_layer"]
        self.final_norm, self.head = layers[-2], layers[-1]
        assert type(self.embed).__name__ == "Position_Embedding" and self.blocks, "KVGenerator needs a GPT layer list"
        self.context = self.embed.context_length
        self.D = self.embed.embed_dim
        self.use_graph = use_graph
        self.len = 0
        self.k = self.v = None
        self.graph = None
        self.graph_kernels = 0
        self._graph_buffers = None
        self._eager_done = False
        self._gen = self._ptrs = None

    def _param_ptrs(self):
        return tuple(getattr(o, pn).ptr for l in self.layers if hasattr(l, "_slots") for o, pn, *_ in l._slots())

    def _params_moved(self):
        """The captured decode graph holds raw parameter addresses.  Training re-homes parameters (ParamArena,
        LayerArenaMixin, fclayer._pad_out) and recycles the old buffers; replaying against them would silently produce
        garbage.  Cheap check (one counter) before every replay; on a real move the graph is dropped and re-captured."""
        if self._gen == rehome_generation():
            return False
        self._gen = rehome_generation()
        ptrs = self._param_ptrs()
        moved, self._ptrs = ptrs != self._ptrs, ptrs
        return moved

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
        if self.k is None or self.k[0].shape[0] != B:
            self.close()
            self.k = [DeviceArray.zeros((B, self.context, self.D)) for _ in self.blocks]
            self.v = [DeviceArray.zeros((B, self.context, self.D)) for _ in self.blocks]
            self.d_pos = DeviceArray.zeros((1,), np.int32)                      # index of the token the next step consumes
            self.d_tok = DeviceArray.zeros((B,), np.int32)                      # that token
            self.d_hist = DeviceArray.zeros((B, self.context + 1), np.int32)    # greedy choices by position
            self._logits = None
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
        self.d_pos.set(np.array([T0], np.int32))
        return np.asarray(x)[:, -1, :]

    def _enqueue_step(self):
        """One decode step on the device: consumes d_tok at position *d_pos, leaves the logits in self._logits, the
        greedy next token in d_tok / d_hist[:, *d_pos + 1], and advances d_pos."""
        B, D = self.d_tok.shape[0], self.D
        # <= 16 sequences: the whole step (embedding, every block, head, greedy choice, position counter) can be recorded as
        # ONE phase program (csrc/phase_prog.cu) instead of 30 launches.  OPT-IN (NTB200_FUSE=1): one cluster of 16 SMs
        # streams the 39 MB of weights of a poetry3000-size step slower than the per-op GEMVs spread over 48-100 SMs
        # (0.174 vs 0.143 ms per token on B200).
        program = B <= 16 and os.environ.get("NTB200_FUSE_DECODE", os.environ.get("NTB200_FUSE", DECODE_PROGRAM_DEFAULT)) == "1"
        if program:
            lib.nt_fuse_begin()
        try:
            self._enqueue_step_ops(B, D)
        finally:
            if program:
                lib.nt_fuse_end()

    def _enqueue_step_ops(self, B, D):
        x = DeviceArray((B, 1, D), host_dtype=self.embed.text_embedding.host_dtype)
        lib.nt_decode_embed(self.d_tok.ptr, self.embed.text_embedding.params.ptr, self.embed.pos_embedding.params.ptr, x.ptr,
                            B, D, self.context, self.d_pos.ptr,       # token row + position row (PatchEmbed.py:132-138)
                            self.embed.text_embedding.num_embeddings)
        fused = B <= 16 and os.environ.get("NTB200_DECODE_FUSED", "1") != "0"
        for bi, blk in enumerate(self.blocks):
            if fused:       # LayerNorm inside the skinny GEMV, KV append inside the decode attention: 6 launches per block
                hdn = self._ln_linear(x, blk.norm, blk.fc0, ops.ACT_RELU)
                h = blk.fc1._forward(hdn, residual=x)
                qkv = self._ln_linear(h, blk.norm1, blk.qkvfc)     # [B,1,3D]
            else:
                u = blk.norm._forward(x)
                hdn = blk.fc0._forward(u, ops.ACT_RELU)
                h = blk.fc1._forward(hdn, residual=x)
                u1 = blk.norm1._forward(h)
                qkv = blk.qkvfc._forward(u1)
                lib.nt_kv_append_at(qkv.ptr, self.k[bi].ptr, self.v[bi].ptr, B, self.context, D, self.d_pos.ptr)
            att = DeviceArray((B, 1, D), host_dtype=x.host_dtype)
            lib.nt_attention_decode_at(qkv.ptr, self.k[bi].ptr, self.v[bi].ptr, att.ptr, B, self.context, blk.num_h,
                                       D // blk.num_h, self.d_pos.ptr, 1 if fused else 0)
            x = blk.fc_out._forward(att, residual=h)
        if fused and not getattr(self.head, "reluact", False):
            z1 = self._ln_linear(x, self.final_norm, self.head.fc0)
            logits = self.head.fc1._forward(z1)
        else:
            z = self.final_norm._forward(x)
            logits = self.head.forward(z)
        V = logits.size // B
        lib.nt_argmax_next(logits.ptr, V, V, B, self.d_tok.ptr, self.d_hist.ptr, self.context + 1, self.d_pos.ptr)
        lib.nt_increment(self.d_pos.ptr)
        self._logits = logits

    @staticmethod
    def _ln_linear(x, norm, fc, act=ops.ACT_NONE):
        """fc(LayerNorm(x)) for the <= 8 rows of a decode step in one launch (nt_linear_ln_skinny)."""
        K, N = fc.params.shape
        M = x.size // K
        y = DeviceArray(x.shape[:-1] + (N,), host_dtype=x.host_dtype)
        lib.nt_linear_ln_skinny(x.ptr, norm.gamma.ptr, norm.beta.ptr, fc.params.ptr, fc.bias_params.ptr if fc.bias else None,
                                y.ptr, M, K, N, act, None)
        return y

    def _run_step(self):
        """Eager the first time, captured the second, replayed afterwards (same protocol as TrainStep.run_device)."""
        if self.graph is not None and self._params_moved():
            lib.nt_sync()
            lib.nt_graph_destroy(self.graph)                   # parameters live elsewhere now: capture again
            self.graph = self._graph_buffers = None
            self._eager_done = False
        if not self.use_graph or not self._eager_done:
            self._enqueue_step()
            self._eager_done = True
        elif self.graph is None:
            lib.nt_sync()
            with keep_allocations() as keep:
                lib.nt_graph_begin()
                try:
                    self._enqueue_step()
                finally:
                    g = ctypes.c_void_p()
                    nk = ctypes.c_int(0)
                    lib.nt_graph_end(ctypes.byref(g), ctypes.byref(nk))
            self._graph_buffers = keep.buffers
            self.graph, self.graph_kernels = g, nk.value
            self._gen, self._ptrs = rehome_generation(), self._param_ptrs()
            lib.nt_graph_launch(self.graph)
        else:
            lib.nt_graph_launch(self.graph)
        self.len += 1

    def step(self, token):
        """Push one new token per sample ([B] ints) through the stack using the caches; returns its logits [B, V]."""
        assert self.k is not None and self.len < self.context, "prefill first / window is full"
        self.d_tok.set(np.asarray(token).astype(np.int32).reshape(-1))
        self._run_step()
        return self._host_logits()

    def _host_logits(self):
        """[B, V]: a head padded by TrainStep (fclayer._pad_out) carries dummy classes with bias -1e30 behind the vocabulary."""
        fc1 = self.head.fc1
        return np.asarray(self._logits).reshape(self.d_tok.shape[0], -1)[:, :getattr(fc1, "_logical_out", fc1.outfeature)]

    def greedy_steps(self, first_token, n):
        """n cached steps with the arg-max taken on the device (np.argmax tie-break): no host round trip between tokens.
        `first_token` [B] is the token at position self.len; returns the n tokens that follow it, [B, n]."""
        assert self.k is not None and self.len + n <= self.context, "prefill first / window too short for n cached steps"
        p0 = self.len
        self.d_tok.set(np.asarray(first_token).astype(np.int32).reshape(-1))
        for _ in range(n):
            self._run_step()
        return self.d_hist.numpy()[:, p0 + 1:p0 + 1 + n].astype(np.int64)

    def close(self):
        """Destroy the captured graph (it pins the buffers of one batch size)."""
        if self.graph is not None:
            lib.nt_graph_destroy(self.graph)
        self.graph = self._graph_buffers = None
        self._eager_done = False

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def generate(self, prompt, n_new, sample=None):
        """Greedy (sample=None) or sampled (sample(probs [B,V]) -> [B] ints) continuation of prompt [B, T0]."""
        out = [list(map(int, row)) for row in np.asarray(prompt)]
        logits = self.prefill(np.asarray(prompt)[:, -self.context:])
        produced = 0
        while produced < n_new:
            if sample is None:
                nxt = np.argmax(logits, axis=-1)
            else:
                e = np.exp(logits - logits.max(-1, keepdims=True))
                nxt = np.asarray(sample(e / e.sum(-1, keepdims=True)))
            for row, t in zip(out, nxt):
                row.append(int(t))
            produced += 1
            if produced == n_new:
                break
            room = self.context - self.len
            if room > 0 and sample is None:
                # greedy: every remaining cached step in one go; the last one's logits decide the token after them
                n = min(room, n_new - produced)
                toks = self.greedy_steps(nxt, n)
                for row, ts in zip(out, toks[:, :n - 1] if n > 1 else toks[:, :0]):
                    row.extend(int(t) for t in ts)
                produced += n - 1
                logits = self._host_logits()
            elif room > 0:
                logits = self.step(nxt)
            else:                                                  # window full: it slides, every position index changes
                logits = self.full_logits(np.array([row[-self.context:] for row in out]))
        return np.array(out)
