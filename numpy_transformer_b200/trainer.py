"""Whole-step driver over a reference-style layer list.

`TrainStep` runs exactly the loop of transformer_of_image.py:141-159 /
gpt/gpt_train_potry3000.py:256-288 — forward every layer, cross_entropy_loss, then
backward/update/setzero in reverse — but

  * parameters, gradients and Adam* state are re-homed into flat arenas so update+setzero is ONE
    kernel over all parameters (and the DP all-reduce is one/few NCCL calls over the grad arena);
  * the step is captured once into a CUDA graph and replayed (the shapes are tiny and launch-bound);
  * inputs arrive from pinned host staging buffers, loss/argmax go back the same way, so the public
    call `step(host_images, host_labels) -> loss` is an honest end-to-end measurement.

The update is applied after the whole backward instead of layer by layer; a layer's backward never
reads another layer's parameters, so the result is identical to the reference's interleaving.
"""
import ctypes
import numpy as np

from . import install
from ._lib import lib
from .device import DeviceArray, PinnedBuffer, as_device, init, keep_allocations
from . import ops
from . import fused_vit


def _import_layers():
    install()
    from net.layernorm import layer_norm
    from net.flatten import flatten_layer
    from attention import attention_layer
    from classify import classify_layer
    from PatchEmbed import PatchEmbed_flatten, Position_Embedding
    from Position_add import Position_learnable
    from gpt.attdecoderblock import attdecoderblock_layer
    return dict(layer_norm=layer_norm, flatten_layer=flatten_layer, attention_layer=attention_layer,
                classify_layer=classify_layer, PatchEmbed_flatten=PatchEmbed_flatten,
                Position_Embedding=Position_Embedding, Position_learnable=Position_learnable,
                attdecoderblock_layer=attdecoderblock_layer)


def build_vit_layers(batch=100, channels=1, height=28, width=28, n_patch=7, embed_dim=96, heads=4, blocks=6,
                     classes=10, seed=None):
    """Layer list of transformer_of_image.py:56-77 (patch_convolu=0, patchnorm, fixed=1, cls_token=0)."""
    L = _import_layers()
    if seed is not None:
        np.random.seed(seed)
    layers = [L["PatchEmbed_flatten"](embed_dim, (batch, channels, height, width), n_patch, patchnorm=True),
              L["Position_learnable"](n_patch, embed_dim, fixed=1)]
    layers += [L["attention_layer"](embed_dim, heads) for _ in range(blocks)]
    layers += [L["layer_norm"](embed_dim), L["flatten_layer"](),
               L["classify_layer"](embed_dim, batch, n_patch, classes, 0)]
    return layers


def build_gpt_layers(batch=64, context=256, embed_dim=384, heads=6, blocks=5, vocab=2350, adam=True, float32=False,
                     seed=None):
    """Layer list of gpt/gpt_train_potry3000.py:176-207 (per-token head, relu=False)."""
    L = _import_layers()
    if seed is not None:
        np.random.seed(seed)
    kw = dict(adam=adam, float32=float32)
    layers = [L["Position_Embedding"](context, vocab, embed_dim, **kw)]
    layers += [L["attdecoderblock_layer"](embed_dim, heads, **kw) for _ in range(blocks)]
    layers += [L["layer_norm"](embed_dim, **kw),
               L["classify_layer"](embed_dim, batch, 1, vocab, cls_token=False, relu=False, **kw)]
    return layers


def save_walk(layers):
    """Checkpoint walk of transformer_of_image.py:200-208."""
    return [l.save_model() for l in layers if 'save_model' in dir(l) and 'restore_model' in dir(l)]


def restore_walk(layers, tree):
    """Restore walk of transformer_of_image.py:84-92."""
    it = iter(tree)
    for l in layers:
        if 'save_model' in dir(l) and 'restore_model' in dir(l):
            l.restore_model(next(it))


class ParamArena:
    """Flat fp32 arenas for parameters / gradients / Adam* moments; layer attributes become views."""

    def __init__(self, layers, adam):
        slots = []
        for l in layers:
            if hasattr(l, "_slots"):
                slots.extend(l._slots())
        self.slots = slots
        offs, n = [], 0
        for obj, pn, gn, mn, vn in slots:
            offs.append(n)
            n += (getattr(obj, pn).size + 3) // 4 * 4           # 16-byte aligned tensors
        self.offsets, self.n = offs, n
        self.params = DeviceArray.zeros((n,))
        self.grads = DeviceArray.zeros((n,))
        self.m = DeviceArray.zeros((n,)) if adam else None
        self.v = DeviceArray.zeros((n,)) if adam else None
        for (obj, pn, gn, mn, vn), off in zip(slots, offs):
            p = getattr(obj, pn)
            g = getattr(obj, gn)
            pv = self.params.view(off, p.shape)
            pv.host_dtype = p.host_dtype
            lib.nt_d2d(pv.ptr, p.ptr, p.nbytes)
            gv = self.grads.view(off, g.shape)
            gv.host_dtype = g.host_dtype
            lib.nt_d2d(gv.ptr, g.ptr, g.nbytes)
            setattr(obj, pn, pv)
            setattr(obj, gn, gv)
            if adam and mn is not None:
                for arena, name in ((self.m, mn), (self.v, vn)):
                    old = getattr(obj, name)
                    nv = arena.view(off, p.shape)
                    if old is not None:
                        lib.nt_d2d(nv.ptr, old.ptr, old.nbytes)
                    setattr(obj, name, nv)
        self.n_params = sum(getattr(o, pn).size for o, pn, *_ in slots)

    def first_offset(self, layer):
        """Arena offset of the first parameter slot owned by `layer` (None if it has no parameters)."""
        if not hasattr(layer, "_slots"):
            return None
        owned = {id(s[0]) for s in layer._slots()}
        for (obj, *_), off in zip(self.slots, self.offsets):
            if id(obj) in owned:
                return off
        return None

    def share_with(self, layers):
        """Point another replica of the same layer list (same constructor order) at these arenas: the
        replica keeps its own activations but reads / accumulates the same parameters and gradients."""
        other = []
        for l in layers:
            if hasattr(l, "_slots"):
                other.extend(l._slots())
        assert len(other) == len(self.slots), "replica has a different parameter structure"
        for (o0, pn, gn, mn, vn), (o1, pn1, gn1, mn1, vn1) in zip(self.slots, other):
            assert getattr(o0, pn).shape == getattr(o1, pn1).shape
            for a in (pn, gn, mn, vn):
                if a is not None and getattr(o0, a, None) is not None:
                    setattr(o1, a, getattr(o0, a))


class TrainStep:
    """One training step of a reference-style layer list as a replayable CUDA graph.

    kind='vit': inputs (B,C,H,W) float images, labels (B,) int   -> N = B        (loss.py:49-50)
    kind='gpt': inputs (B,T) int tokens,      labels (B,T) int   -> N = B*T, causal mask in-kernel
    dp: None or a DataParallel (parallel.py); the local batch is one shard of the global batch and
        the loss/gradient normaliser is the GLOBAL row count.
    """

    def __init__(self, layers, kind, input_shape, lr, adam=False, dp=None, use_graph=True, grad_buckets=1,
                 replicas=()):
        """replicas: extra copies of the layer list (same constructors).  With r replicas the local batch is
        cut into r+1 micro-batches that run as parallel branches of the captured step (intra-GPU data
        parallelism: same parameters, gradients accumulate atomically into the same arena, loss normalised
        by the full row count) — the tiny ViT kernels are latency-bound, so their latencies overlap."""
        init()
        self.layers, self.kind, self.adam, self.dp = layers, kind, adam, dp
        self.use_graph = use_graph
        self.chains = [layers] + [list(r) for r in replicas]
        for ch in self.chains:                  # the head feeds cross_entropy directly: pad its class axis
            head = ch[-1]
            if hasattr(head, "fc1") and hasattr(head.fc1, "_pad_out"):
                head.fc1._pad_out(4)
        self.arena = ParamArena(layers, adam)
        for r in self.chains[1:]:
            self.arena.share_with(r)
        assert input_shape[0] % len(self.chains) == 0, "batch must divide evenly into micro-batch chains"
        self.input_shape = tuple(input_shape)
        in_dtype = np.float32 if kind == "vit" else np.int32
        self.rows = input_shape[0] if kind == "vit" else input_shape[0] * input_shape[1]
        world = 1 if dp is None else dp.world
        self.n_norm = float(self.rows * world)
        self.h_in = PinnedBuffer(self.input_shape, in_dtype)
        self.h_lab = PinnedBuffer((self.rows,), np.int32)
        self.h_out = PinnedBuffer((8,), np.float32)
        self.h_amax = PinnedBuffer((self.rows,), np.int32)
        self.d_in = DeviceArray(self.input_shape, in_dtype)
        self.d_lab = DeviceArray((self.rows,), np.int32)
        self.d_loss = DeviceArray.zeros((len(self.chains),))
        self.d_amax = DeviceArray.zeros((self.rows,), dtype=np.int32)
        self.d_lr = DeviceArray.from_host(np.array([lr], dtype=np.float32)).reshape(())
        self.d_t = DeviceArray.from_host(np.array([1], dtype=np.int32)).reshape(())
        self._lr = float(lr)
        self.graph = None
        self.graph_kernels = 0
        self.steps_done = 0
        self.grad_buckets = max(1, int(grad_buckets))
        self.loss = self.logits = self.probs = self.argmax = None
        L = _import_layers()
        self._block_t = L["attdecoderblock_layer"]
        self._vit_block_t = L["attention_layer"]
        self._layer_types = L

    # ------------------------------------------------------------------ fully fused ViT step (README layer list)
    def _vit_fast_plan(self, layers):
        """The README layer list (transformer_of_image.py:56-77: PatchEmbed_flatten(patchnorm) -> fixed position table ->
        attention blocks -> layer_norm -> flatten -> classify_layer(relu, no cls token)) at shapes the fused kernels
        cover runs as 9 launches: stem, weight split, block stack, 3 x head, block stack backward, stem backward, update."""
        if self.kind != "vit" or len(self.chains) != 1 or len(layers) < 6:
            return None
        L = self._layer_types
        pe, pos, ln, fl, head = layers[0], layers[1], layers[-3], layers[-2], layers[-1]
        blocks = layers[2:-3]
        if not (type(pe) is L["PatchEmbed_flatten"] and type(pos) is L["Position_learnable"] and type(ln) is L["layer_norm"]
                and type(fl) is L["flatten_layer"] and type(head) is L["classify_layer"]):
            return None
        if not (pe.patchnorm and pos.fixed and not head.cls_token and head.reluact and ln.elementwise_affine
                and pe.fullconnect.bias and head.fc0.bias and head.fc1.bias):
            return None
        B, C, Hi, Wi = self.input_shape
        N, D = pe.n_patch ** 2, pe.embed_dim
        if Hi % pe.n_patch or Wi % pe.n_patch or not blocks:
            return None
        if fused_vit.fusable_run(layers, 2, self._vit_block_t, (B, N, D)) != len(blocks):
            return None
        D0, C1 = head.fc0.params.shape[1], head.fc1.params.shape[1]
        if head.fc0.params.shape[0] != N * D or not fused_vit.ends_supported(B, C, Hi, Wi, pe.n_patch, D, D0, C1):
            return None
        return dict(pe=pe, pos=pos, ln=ln, head=head, blocks=blocks, B=B, C=C, Hi=Hi, Wi=Wi, N=N, D=D, D0=D0, C1=C1)

    def _enqueue_vit_fast(self, plan, layers):
        import ctypes
        pe, pos, ln, head, blocks = plan["pe"], plan["pos"], plan["ln"], plan["head"], plan["blocks"]
        B, C, Hi, Wi, N, D, D0, C1 = (plan[k] for k in ("B", "C", "Hi", "Wi", "N", "D", "D0", "C1"))
        if getattr(self, "_ends_h0", None) is None:
            self._ends_h0 = DeviceArray.zeros((B, D0))           # fc0 accumulator: zero between steps (the kernel re-zeroes it)
        e = fused_vit.NtVitEnds()
        x0, pe_lin, pe_stat = DeviceArray((B, N, D)), DeviceArray((B, N, D)), DeviceArray((B, N, 2))
        tl_stat, dh0, dxl = DeviceArray((B, N, 2)), DeviceArray((B, D0)), DeviceArray((B, N, D))
        logits, probs = DeviceArray((B, C1)), DeviceArray((B, C1))
        fc, n0 = pe.fullconnect, pe.norm
        e.images = self.d_in.ptr
        e.wpe, e.bpe, e.pe_g, e.pe_b = fc.params.ptr, fc.bias_params.ptr, n0.gamma.ptr, n0.beta.ptr
        e.pos = pos.table().ptr
        e.d_wpe, e.d_bpe, e.d_pe_g, e.d_pe_b = fc.params_delta.ptr, fc.bias_delta.ptr, n0.gamma_delta.ptr, n0.beta_delta.ptr
        e.pe_lin, e.pe_stat, e.x0 = pe_lin.ptr, pe_stat.ptr, x0.ptr
        e.tl_g, e.tl_b = ln.gamma.ptr, ln.beta.ptr
        e.w0, e.b0, e.w1, e.b1 = head.fc0.params.ptr, head.fc0.bias_params.ptr, head.fc1.params.ptr, head.fc1.bias_params.ptr
        e.d_tl_g, e.d_tl_b = ln.gamma_delta.ptr, ln.beta_delta.ptr
        e.d_w0, e.d_b0 = head.fc0.params_delta.ptr, head.fc0.bias_delta.ptr
        e.d_w1, e.d_b1 = head.fc1.params_delta.ptr, head.fc1.bias_delta.ptr
        e.labels = self.d_lab.ptr
        e.tl_stat, e.h0, e.dh0 = tl_stat.ptr, self._ends_h0.ptr, dh0.ptr
        e.logits, e.probs, e.argmax, e.loss = logits.ptr, probs.ptr, self.d_amax.ptr, self.d_loss.ptr
        e.dxl = dxl.ptr
        n0._lead = ln._lead = 2                                  # gamma/beta host shape after a forward (layernorm.py:105-108)
        lib.nt_vit_stem_fwd(ctypes.addressof(e), B, C, Hi, Wi, pe.n_patch, D)
        xl = fused_vit.forward_stack(blocks, x0)
        e.xl = xl.ptr
        lib.nt_vit_head(ctypes.addressof(e), B, N, D, D0, C1, self.n_norm)
        if self.dp is not None and self.dp.world > 1:
            # gradients of the final LayerNorm + head (the tail of the arena) are complete: all-reduce them under the
            # fused backward launch
            lo = min(o for o in (self.arena.first_offset(ln), self.arena.first_offset(head)) if o is not None)
            lib.nt_dp_allreduce_sum(self.arena.grads.ptr + 4 * lo, self.arena.n - lo)
            self._early_from = lo
        dx0 = fused_vit.backward_stack(blocks, dxl)
        lib.nt_vit_stem_bwd(ctypes.addressof(e), dx0.ptr, B, C, Hi, Wi, pe.n_patch, D)
        probs._argmax = self.d_amax
        self._ends_keep = (e, x0, pe_lin, pe_stat, tl_stat, dh0, dxl, logits, probs)
        return logits, probs

    # ------------------------------------------------------------------ the step body (enqueue only)
    def _enqueue_chain(self, c, layers):
        plan = self._vit_fast_plan(layers)
        if plan is not None:
            return self._enqueue_vit_fast(plan, layers)
        nc = len(self.chains)
        bs = self.input_shape[0] // nc
        rows = self.rows // nc
        in_elems = self.d_in.size // nc
        x = self.d_in.view(c * in_elems, (bs,) + self.input_shape[1:])
        lab = self.d_lab.view(c * rows, (rows,))
        runs = {}                                    # start index -> fused run of ViT blocks (one launch each way)
        early = {}     # runs whose weights were split ahead of time (fused_vit.prepare); measured: forking the split onto
        # the side branch costs more in graph dependencies than the 5 us launch it hides, so it stays in the forward call
        i = 0
        while i < len(layers):
            l = layers[i]
            n = fused_vit.fusable_run(layers, i, self._vit_block_t, x.shape) if type(l) is self._vit_block_t else 0
            if n:
                runs[i] = layers[i:i + n]
                prepared = early.get(i) == n
                if prepared:
                    lib.nt_side_join()
                x = fused_vit.forward_stack(runs[i], x, prepared=prepared)
                i += n
                continue
            if isinstance(l, self._block_t):
                x = l.forward(x, "causal")
            else:
                x = l.forward(x)
            i += 1
        flat = x.reshape(rows, x.shape[-1])
        loss, grad, probs, amax = ops.cross_entropy(flat, lab, None, self.n_norm, loss_out=self.d_loss.view(c, ()),
                                                    argmax_out=self.d_amax.view(c * rows, (rows,)))
        delta = grad.reshape(x.shape)
        first = layers[0]
        ends = {s + len(r) - 1: s for s, r in runs.items()}
        i = len(layers) - 1
        while i >= 0:
            l = layers[i]
            if i in ends:
                if self.dp is not None and self.dp.world > 1 and len(self.chains) == 1 and self._early_from is None:
                    # data parallel: the gradients of everything AFTER the fused stack (final LayerNorm + head, more than
                    # half of the V1 arena) are complete -> all-reduce them now, under the fused backward launch
                    offs = [self.arena.first_offset(l) for l in layers[i + 1:]]
                    offs = [o for o in offs if o is not None]
                    if offs:
                        lib.nt_side_join()
                        lo = min(offs)
                        lib.nt_dp_allreduce_sum(self.arena.grads.ptr + 4 * lo, self.arena.n - lo)
                        self._early_from = lo
                delta = fused_vit.backward_stack(runs[ends[i]], delta)
                i = ends[i] - 1
                continue
            if l is first and (hasattr(l, "fullconnect") or hasattr(l, "convolution")):
                delta = l.backward(delta, want_dx=False)     # nobody consumes the patch-space gradient
            else:
                delta = l.backward(delta)
            i -= 1
        return x, probs

    def _enqueue(self):
        nc = len(self.chains)
        outs = []
        self._early_from = None          # arena offset from which the gradients were all-reduced early (DP only)
        for c, layers in enumerate(self.chains):
            if nc > 1:
                lib.nt_chain_begin(c)
            outs.append(self._enqueue_chain(c, layers))
            if nc > 1:
                lib.nt_chain_end()
        if nc > 1:
            lib.nt_chain_join()
        self.logits, self.probs = outs[0] if nc == 1 else ([o[0] for o in outs], [o[1] for o in outs])
        self.loss, self.argmax = self.d_loss, self.d_amax
        a = self.arena
        lib.nt_side_join()                       # weight-gradient branch must land before all-reduce / update
        if self.dp is not None and self.dp.world > 1:
            if self._early_from is None:
                self.dp.allreduce_buckets(a.grads, self.grad_buckets)
            elif self._early_from > 0:
                self.dp.allreduce_buckets(a.grads.view(0, (self._early_from,)), self.grad_buckets)
            lib.nt_dp_join()
        if self.adam:
            lib.nt_adam_star(a.params.ptr, a.grads.ptr, a.m.ptr, a.v.ptr, a.n, 0.0, self.d_lr.ptr, 0, self.d_t.ptr, 1)
            lib.nt_increment(self.d_t.ptr)
        else:
            lib.nt_sgd(a.params.ptr, a.grads.ptr, a.n, 0.0, self.d_lr.ptr, 1)

    def set_lr(self, lr):
        if float(lr) != self._lr:
            self._lr = float(lr)
            self.d_lr.set(np.array([lr], dtype=np.float32))

    def run_device(self):
        """One step on whatever is in the device input buffers (no host traffic).  Call 1 runs
        eagerly (sizes the pool, loads modules), call 2 captures the graph and launches it, later
        calls replay it; every call is exactly one training step."""
        if not self.use_graph or self.steps_done == 0:
            self._enqueue()
        elif self.graph is None:
            lib.nt_sync()
            with keep_allocations() as keep:
                lib.nt_graph_begin()
                try:
                    self._enqueue()
                finally:
                    g = ctypes.c_void_p()
                    nk = ctypes.c_int(0)
                    lib.nt_graph_end(ctypes.byref(g), ctypes.byref(nk))
            self._graph_buffers = keep.buffers
            self.graph, self.graph_kernels = g, nk.value
            lib.nt_graph_launch(self.graph)
        else:
            lib.nt_graph_launch(self.graph)
        self.steps_done += 1

    def capture_profiled(self):
        """Capture the step once more with an (external) CUDA-event pair around every C-ABI op.  Returns
        (graph, records); after `lib.nt_graph_launch(graph)` + sync, `lib.profile_read(records)` yields the
        per-op device durations of that replay.  Every replay is a real training step."""
        lib.nt_sync()
        lib.nt_set_side_branch(0)               # serial capture: per-op times free of side-branch interference
        with keep_allocations() as keep:
            lib.profile_begin(external=True)
            lib.nt_graph_begin()
            try:
                self._enqueue()
            finally:
                recs = lib.profile_pause()
                g = ctypes.c_void_p()
                nk = ctypes.c_int(0)
                lib.nt_graph_end(ctypes.byref(g), ctypes.byref(nk))
        lib.nt_set_side_branch(1)
        self._prof_buffers = keep.buffers
        return g, recs

    def load(self, inputs=None, labels=None):
        """Host -> pinned staging -> device (async on the compute stream).  With inputs / labels omitted the batch
        is taken from the pinned staging buffers as they are (`self.h_in.array`, `self.h_lab.array`): a loader that
        writes its batches straight into pinned memory skips the extra host copy / dtype cast."""
        if inputs is not None:
            np.copyto(self.h_in.array, np.asarray(inputs).reshape(self.input_shape), casting="unsafe")
        if labels is not None:
            np.copyto(self.h_lab.array, np.asarray(labels).reshape(-1), casting="unsafe")
        lib.nt_h2d(self.d_in.ptr, self.h_in.ptr, self.h_in.nbytes)
        lib.nt_h2d(self.d_lab.ptr, self.h_lab.ptr, self.h_lab.nbytes)

    def step(self, inputs=None, labels=None, lr=None, want_argmax=False):
        """End-to-end public call: host batch in (or already in the pinned staging buffers), host loss (and argmax) out."""
        if lr is not None:
            self.set_lr(lr)
        self.load(inputs, labels)
        self.run_device()
        nc = len(self.chains)
        lib.nt_d2h_async(self.h_out.ptr, self.d_loss.ptr, 4 * nc)
        if want_argmax:
            lib.nt_d2h_async(self.h_amax.ptr, self.d_amax.ptr, self.h_amax.nbytes)
        lib.nt_sync()
        loss = float(self.h_out.array[:nc].sum())      # each chain's partial is already divided by the global N
        return (loss, self.h_amax.array.copy()) if want_argmax else loss

    @property
    def h2d_bytes(self):
        return self.h_in.nbytes + self.h_lab.nbytes

    def sync_layer_counters(self):
        """Bring the per-object Adam `t` (fullconnect.py:97,149) in line with the device counter."""
        t = int(self.d_t)
        for obj, *_ in self.arena.slots:
            obj.t = t

    # ------------------------------------------------------------------ resumable checkpoints
    def state_dict(self):
        """Everything needed to resume bit-for-bit: the reference's own checkpoint tree (`save_walk`, what its
        trainers pickle: transformer_of_image.py:200-208) PLUS the optimizer state the reference never saves
        (Adam* moments and step counter, net/fullconnect.py:66-73,97) and the current learning rate."""
        self.sync_layer_counters()
        out = {"model": save_walk(self.layers), "lr": self._lr, "steps_done": self.steps_done, "adam": bool(self.adam),
               "t": int(self.d_t)}
        if self.adam:
            out["adam_m"] = self.arena.m.numpy().copy()
            out["adam_v"] = self.arena.v.numpy().copy()
            out["arena_size"] = self.arena.n
        return out

    def load_state_dict(self, state):
        """Inverse of state_dict().  A plain reference checkpoint (the nested list alone) is accepted too: parameters are
        restored and the optimizer state starts fresh, exactly what the reference's own restore does."""
        if not isinstance(state, dict):
            state = {"model": state}
        restore_walk(self.layers, state["model"])
        if "lr" in state:
            self.set_lr(state["lr"])
        if self.adam and "adam_m" in state:
            assert state.get("arena_size") == self.arena.n, "optimizer state belongs to a different model"
            self.arena.m.set(state["adam_m"])
            self.arena.v.set(state["adam_v"])
            self.d_t.set(np.array([state["t"]], dtype=np.int32))
            self.sync_layer_counters()
        lib.nt_sync()

    def close(self):
        """Destroy the captured graphs (required before DataParallel.shutdown: they hold NCCL kernels)."""
        lib.nt_sync()
        if self.graph is not None:
            lib.nt_graph_destroy(self.graph)
            self.graph = None
        self._graph_buffers = None

    def __del__(self):
        try:
            if self.graph is not None:
                lib.nt_graph_destroy(self.graph)
        except Exception:
            pass
