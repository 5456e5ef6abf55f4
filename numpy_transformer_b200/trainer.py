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
import os
import numpy as np

from . import install
from ._lib import lib
from .device import DeviceArray, PinnedBuffer, as_device, init, keep_allocations, rehome_generation, bump_rehome
from . import ops
from . import fused_vit
from .parallel import bucket_bounds


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


def restore_walk(layers, tree, grow=False):
    """Restore walk of transformer_of_image.py:84-92.  grow=True is the progressive-growth restore of
    gpt/gpt_train_potry3000.py:212-231: the checkpoint was written by a model with FEWER decoder blocks; a layer whose
    entry does not fit (the new block meets the final LayerNorm's [gamma, beta]) is seeded from the PREVIOUS entry — the
    block before it — and the walk continues without consuming an entry.  A trailing int (the step counter `jk` the
    trainer appends, :370) is returned; -> (step counter or None, number of layers seeded from their predecessor)."""
    tree = list(tree)
    jk = tree[-1] if tree and isinstance(tree[-1], (int, np.integer)) else None
    cnt = grown = 0
    for l in layers:
        if 'save_model' in dir(l) and 'restore_model' in dir(l):
            if not grow:
                l.restore_model(tree[cnt])
            else:
                try:
                    l.restore_model(tree[cnt])
                except (IndexError, ValueError, TypeError):
                    l.restore_model(tree[cnt - 1])
                    grown += 1
                    continue
            cnt += 1
    return (None if grown else jk), grown


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
        self.adam = adam
        bump_rehome()

    def readopt(self):
        """Bring back every tensor that has left this arena since it was built (a layer re-homed into its own arena by
        a hand-made layer.update(), a head padded by another TrainStep ...): its current values are copied into the
        arena slot and the attribute points at the slot again, so a graph captured over the arena stays valid and
        save_walk() sees what the graph trains.  -> number of tensors moved back."""
        moved = 0
        for (obj, pn, gn, mn, vn), off in zip(self.slots, self.offsets):
            names = [(pn, self.params), (gn, self.grads)]
            if self.adam and mn is not None:
                names += [(mn, self.m), (vn, self.v)]
            for name, arena in names:
                cur = getattr(obj, name)
                want = arena.ptr + 4 * off
                if cur is None or cur.ptr == want:
                    continue
                view = arena.view(off, cur.shape)
                view.host_dtype = cur.host_dtype
                assert view.size == cur.size, "parameter %s changed shape after the arena was built" % name
                lib.nt_d2d(view.ptr, cur.ptr, cur.nbytes)
                setattr(obj, name, view)
                moved += 1
        return moved

    def first_offset(self, layer):
        """Arena offset of the first parameter slot owned by `layer` (None if it has no parameters)."""
        if not hasattr(layer, "_slots"):
            return None
        owned = {id(s[0]) for s in layer._slots()}
        for (obj, *_), off in zip(self.slots, self.offsets):
            if id(obj) in owned:
                return off
        return None

    def optimizer_runs(self, adam):
        """[(is_adam, first float, count)]: maximal runs of consecutive slots updated by the same rule.  A slot is
        Adam* when the step is adam AND the layer declares moments for it AND the owning layer was built with adam
        (fullconnect.py:134-153); everything else is plain SGD (e.g. Position_learnable.param)."""
        runs = []
        for (obj, pn, gn, mn, vn), off in zip(self.slots, self.offsets):
            size = (getattr(obj, pn).size + 3) // 4 * 4
            is_adam = bool(adam and mn is not None and getattr(obj, "adam", True))
            if runs and runs[-1][0] == is_adam and runs[-1][1] + runs[-1][2] == off:
                runs[-1][2] += size
            else:
                runs.append([is_adam, off, size])
        return [tuple(r) for r in runs]


class TrainStep:
    """One training step of a reference-style layer list as a replayable CUDA graph.

    kind='vit': inputs (B,C,H,W) float images, labels (B,) int   -> N = B        (loss.py:49-50)
    kind='gpt': inputs (B,T) int tokens,      labels (B,T) int   -> N = B*T, causal mask in-kernel
    dp: None or a DataParallel (parallel.py); the local batch is one shard of the global batch and
        the loss/gradient normaliser is the GLOBAL row count.
    """

    def __init__(self, layers, kind, input_shape, lr, adam=False, dp=None, use_graph=True, grad_buckets=1):
        init()
        self.layers, self.kind, self.adam, self.dp = layers, kind, adam, dp
        self.use_graph = use_graph
        head = layers[-1]                       # the head feeds cross_entropy directly: pad its class axis
        if hasattr(head, "fc1") and hasattr(head.fc1, "_pad_out"):
            head.fc1._pad_out(4)
        self.arena = ParamArena(layers, adam)
        self.input_shape = tuple(input_shape)
        in_dtype = np.float32 if kind == "vit" else np.int32
        self.rows = input_shape[0] if kind == "vit" else input_shape[0] * input_shape[1]
        world = 1 if dp is None else dp.world
        self.n_norm = float(self.rows * world)
        # two pinned staging sets: load() fills one while the H2D copy of the previous batch may still be reading the
        # other (pipelined load()/run_device() without a sync in between)
        self._stage = [(PinnedBuffer(self.input_shape, in_dtype), PinnedBuffer((self.rows,), np.int32), self._new_event())
                       for _ in range(2)]
        self._stage_i = 0
        self._stage_used = [False, False]
        self._e2e_graphs = {}
        self.h_out = PinnedBuffer((8,), np.float32)
        self.h_amax = PinnedBuffer((self.rows,), np.int32)
        self.d_in = DeviceArray(self.input_shape, in_dtype)
        self.d_lab = DeviceArray((self.rows,), np.int32)
        self.d_loss = DeviceArray.zeros((1,))
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
        self._opt_runs = self.arena.optimizer_runs(adam)
        # launch-bound GPT shapes (<= 16 token rows: the gpt_character configuration): forward, loss and backward can be
        # recorded as ONE phase program (csrc/phase_prog.cu, 3 launches per step instead of 33).  OPT-IN (NTB200_FUSE=1):
        # measured on B200 the program is slower than the graph of per-op launches it replaces (0.184 vs 0.123 ms per
        # gpt_character step) — a dependent phase costs ~5.8 us through the cluster barrier against ~4 us
        # per kernel under programmatic dependent launch.
        self._fuse = kind == "gpt" and self.rows <= 16 and os.environ.get("NTB200_FUSE", "0") == "1"
        if dp is not None and dp.world > 1 and dp.nccl_ready:
            dp.init_p2p(self.arena.grads)           # small arenas: one-kernel NVLink all-reduce instead of an NCCL call
        self._gen = rehome_generation()
        self.sync_replicas()

    @staticmethod
    def _new_event():
        e = ctypes.c_void_p()
        lib.nt_event_create(ctypes.byref(e))
        return e

    @property
    def h_in(self):
        """The pinned input staging buffer the NEXT load() / step() without arguments will upload."""
        return self._stage[self._stage_i][0]

    @property
    def h_lab(self):
        return self._stage[self._stage_i][1]

    def sync_replicas(self):
        """Data parallel: every rank must hold rank 0's parameters and optimizer state (the reference has one model;
        ranks that seeded np.random differently, or restored different checkpoints, would otherwise train different
        models without any error).  Broadcast the arenas from rank 0, then verify with a checksum all-reduce."""
        dp = self.dp
        if dp is None or dp.world == 1 or not dp.nccl_ready:
            return
        a = self.arena
        bufs = [a.params] + ([a.m, a.v] if self.adam else [])
        for b in bufs:
            lib.nt_dp_broadcast(b.ptr, b.size, 0)
        lib.nt_dp_broadcast(self.d_t.ptr, 1, 0)            # int32 moved as one 4-byte word
        lib.nt_sync()
        mine = float(np.asarray(a.params.numpy(), dtype=np.float64).sum())
        if abs(dp.max_over_ranks(mine) + dp.max_over_ranks(-mine)) > 0:
            raise RuntimeError("data-parallel replicas differ after the parameter broadcast")

    # ------------------------------------------------------------------ the step body (enqueue only)
    def _enqueue_body(self):
        layers = self.layers
        x, lab, rows = self.d_in, self.d_lab, self.rows
        runs = {}                                    # start index -> fused run of ViT blocks (one launch each way)
        i = 0
        while i < len(layers):
            l = layers[i]
            n = fused_vit.fusable_run(layers, i, self._vit_block_t, x.shape) if type(l) is self._vit_block_t else 0
            if n:
                runs[i] = layers[i:i + n]
                x = fused_vit.forward_stack(runs[i], x)
                i += n
                continue
            if isinstance(l, self._block_t):
                x = l.forward(x, "causal")
            else:
                x = l.forward(x)
            i += 1
        flat = x.reshape(rows, x.shape[-1])
        loss, grad, probs, amax = ops.cross_entropy(flat, lab, None, self.n_norm, loss_out=self.d_loss.view(0, ()),
                                                    argmax_out=self.d_amax)
        delta = grad.reshape(x.shape)
        first = layers[0]
        ends = {s + len(r) - 1: s for s, r in runs.items()}
        i = len(layers) - 1
        while i >= 0:
            l = layers[i]
            if i in ends:
                if self.dp is not None and self.dp.world > 1 and self._early_from is None and os.environ.get("NTB200_DP_SKIP") != "1":
                    # data parallel: the gradients of everything AFTER the fused stack (final LayerNorm + head, more than
                    # half of the V1 arena) are complete -> all-reduce them now, under the fused backward launch
                    offs = [self.arena.first_offset(l) for l in layers[i + 1:]]
                    offs = [o for o in offs if o is not None]
                    if offs:
                        lib.nt_side_join()
                        lo = min(offs)
                        self.dp.allreduce_range(self.arena.grads, lo, self.arena.n, side=True)
                        self._early_from = lo
                run = runs[ends[i]]
                dp_on = (self.dp is not None and self.dp.world > 1 and self._early_from is not None
                         and os.environ.get("NTB200_DP_SKIP") != "1")
                delta = fused_vit.backward_stack(run, delta)
                # With the one-kernel NVLink exchange available (small arenas) the block gradients wait for the stem's and
                # go in ONE exchange at the end of the step: an exchange costs >= 20 us at 8 GPUs whatever its size, and the
                # stem's backward (~10 us) hides less of a separate block exchange than the second latency floor costs
                # (8 GPUs: 0.5359 vs 0.5430 ms/step).  NTB200_DP_MERGE=0 restores the early block exchange.
                merge = getattr(self.dp, "p2p_ready", False) and os.environ.get("NTB200_DP_MERGE", "1") != "0"
                if dp_on and not merge:
                    # the block gradients are complete: reduce them NOW on the comm stream, under the stem's backward
                    # kernels; only the (tiny) range in front of the blocks is left for the end of the step
                    lo = self.arena.first_offset(layers[ends[i]])
                    if lo is not None and lo < self._early_from:
                        self.dp.allreduce_range(self.arena.grads, lo, self._early_from, side=True)
                        self._early_from = lo
                i = ends[i] - 1
                continue
            if l is first and (hasattr(l, "fullconnect") or hasattr(l, "convolution")):
                delta = l.backward(delta, want_dx=False)     # nobody consumes the patch-space gradient
            else:
                delta = l.backward(delta)
            i -= 1
        return x, probs

    def _enqueue(self):
        self._early_from = None          # arena offset from which the gradients were all-reduced early (DP only)
        if self._fuse:
            lib.nt_fuse_begin()
        try:
            with ops.lean_attention():       # nobody reads the attention probabilities of a TrainStep (they are recomputed on demand)
                self.logits, self.probs = self._enqueue_body()
        finally:
            if self._fuse:
                lib.nt_fuse_end()            # launches the recorded program
        self.loss, self.argmax = self.d_loss, self.d_amax
        a = self.arena
        lib.nt_side_join()                       # weight-gradient branch must land before all-reduce / update
        if self.dp is not None and self.dp.world > 1 and os.environ.get("NTB200_DP_SKIP") == "1":
            pass            # measurement aid: replicas run unsynchronised (what the step costs without any exchange)
        elif self.dp is not None and self.dp.world > 1:
            hi = a.n if self._early_from is None else self._early_from
            if hi > 0:
                if getattr(self.dp, "p2p_ready", False) and hi <= self.dp.p2p_max_floats:
                    # launch-bound shapes: ONE NVLink peer-memory kernel on the compute stream right behind the backward
                    self.dp.allreduce_range(a.grads, 0, hi, side=False)
                else:
                    for lo_, hi_ in bucket_bounds(hi, self.grad_buckets):
                        self.dp.allreduce_range(a.grads, lo_, hi_, side=True)
            lib.nt_dp_join()
        # one launch per run of slots that share an optimizer: Adam* where the layer has moments and adam=True, plain SGD
        # elsewhere (Position_learnable.param has no moments in the reference: Position_add.py:26-28 is always SGD)
        for is_adam, lo, n in self._opt_runs:
            if is_adam:
                lib.nt_adam_star(a.params.ptr + 4 * lo, a.grads.ptr + 4 * lo, a.m.ptr + 4 * lo, a.v.ptr + 4 * lo, n, 0.0,
                                 self.d_lr.ptr, 0, self.d_t.ptr, 1)
            else:
                lib.nt_sgd(a.params.ptr + 4 * lo, a.grads.ptr + 4 * lo, n, 0.0, self.d_lr.ptr, 1)
        if self.adam:
            lib.nt_increment(self.d_t.ptr)

    def device_barrier(self):
        """Data parallel: the ranks' compute streams meet here — a 16-byte all-reduce of gradient words that are zero
        between steps.  bench.py puts it between the L2 flush and the timed step so that one rank's slower flush is not
        charged to the others' step (they would wait for it inside the gradient exchange)."""
        if self.dp is not None and self.dp.world > 1:
            self.dp.allreduce_range(self.arena.grads, 0, 4, side=False)
            lib.nt_dp_join()

    def set_lr(self, lr):
        if float(lr) != self._lr:
            self._lr = float(lr)
            self.d_lr.set(np.array([lr], dtype=np.float32))

    def run_device(self):
        """One step on whatever is in the device input buffers (no host traffic).  Call 1 runs
        eagerly (sizes the pool, loads modules), call 2 captures the graph and launches it, later
        calls replay it; every call is exactly one training step."""
        if rehome_generation() != self._gen:
            # some parameter tensor moved since the arena was built / the graph captured (generate.py and the layers'
            # own update() re-home tensors): pull strays back so the captured addresses stay the live ones
            self.arena.readopt()
            self._gen = rehome_generation()
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
        fuse, self._fuse = self._fuse, False    # per-op times need per-op launches
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
        self._fuse = fuse
        self._prof_buffers = keep.buffers
        return g, recs

    def load(self, inputs=None, labels=None):
        """Host -> pinned staging -> device (async on the compute stream).  With inputs / labels omitted the batch
        is taken from the current pinned staging buffers as they are (`self.h_in.array`, `self.h_lab.array`): a loader
        that writes its batches straight into pinned memory skips the extra host copy / dtype cast.  Staging is
        double-buffered: the set being filled is never the one an in-flight H2D copy is still reading, so
        load() / run_device() can be called back to back without a sync."""
        i = self._stage_i
        h_in, h_lab, ev = self._stage[i]
        if self._stage_used[i]:
            lib.nt_event_synchronize(ev)          # the copy issued from this set two loads ago has finished
        if inputs is not None:
            np.copyto(h_in.array, np.asarray(inputs).reshape(self.input_shape), casting="unsafe")
        if labels is not None:
            np.copyto(h_lab.array, np.asarray(labels).reshape(-1), casting="unsafe")
        lib.nt_h2d(self.d_in.ptr, h_in.ptr, h_in.nbytes)
        lib.nt_h2d(self.d_lab.ptr, h_lab.ptr, h_lab.nbytes)
        lib.nt_event_record(ev)
        self._stage_used[i] = True
        if inputs is not None or labels is not None:
            self._stage_i = 1 - i                 # a loader that fills h_in itself keeps addressing the same set

    def _step_one_graph(self, inputs, labels):
        """The end-to-end step as ONE graph launch: the H2D copies of the batch out of a pinned staging set, the training step
        and the D2H copies of loss and arg-max are nodes of the same graph (three stream operations and their gaps less per
        call than load() + run_device() + nt_d2h_async: bench.py's e2e 0.528 -> 0.51x ms on the README ViT).  One graph per
        staging set, captured on first use; step() is synchronous, so a set is never overwritten under a running copy."""
        i = self._stage_i
        h_in, h_lab, ev = self._stage[i]
        if self._stage_used[i]:
            lib.nt_event_synchronize(ev)
            self._stage_used[i] = False
        if inputs is not None:
            np.copyto(h_in.array, np.asarray(inputs).reshape(self.input_shape), casting="unsafe")
        if labels is not None:
            np.copyto(h_lab.array, np.asarray(labels).reshape(-1), casting="unsafe")
        if rehome_generation() != self._gen:
            self.arena.readopt()
            self._gen = rehome_generation()
            self._e2e_graphs = {}
        g = self._e2e_graphs.get(i)
        if g is None:
            lib.nt_sync()
            with keep_allocations() as keep:
                lib.nt_graph_begin()
                try:
                    lib.nt_h2d(self.d_in.ptr, h_in.ptr, h_in.nbytes)
                    lib.nt_h2d(self.d_lab.ptr, h_lab.ptr, h_lab.nbytes)
                    self._enqueue()
                    lib.nt_d2h_async(self.h_out.ptr, self.d_loss.ptr, 4)
                    lib.nt_d2h_async(self.h_amax.ptr, self.d_amax.ptr, self.h_amax.nbytes)
                finally:
                    gg = ctypes.c_void_p()
                    nk = ctypes.c_int(0)
                    lib.nt_graph_end(ctypes.byref(gg), ctypes.byref(nk))
            self._e2e_buffers = getattr(self, "_e2e_buffers", []) + [keep.buffers]
            g = self._e2e_graphs[i] = gg
        lib.nt_graph_launch(g)
        self.steps_done += 1
        if inputs is not None or labels is not None:
            self._stage_i = 1 - i

    def step(self, inputs=None, labels=None, lr=None, want_argmax=False):
        """End-to-end public call: host batch in (or already in the pinned staging buffers), host loss (and argmax) out."""
        if lr is not None:
            self.set_lr(lr)
        if (self.use_graph and self.graph is not None and os.environ.get("NTB200_E2E_GRAPH", "1") != "0"
                and (self.dp is None or self.dp.world == 1)):
            self._step_one_graph(inputs, labels)
            lib.nt_sync()
            loss = float(self.h_out.array[0])
            return (loss, self.h_amax.array.copy()) if want_argmax else loss
        self.load(inputs, labels)
        self.run_device()
        lib.nt_d2h_async(self.h_out.ptr, self.d_loss.ptr, 4)
        if want_argmax:
            lib.nt_d2h_async(self.h_amax.ptr, self.d_amax.ptr, self.h_amax.nbytes)
        lib.nt_sync()
        loss = float(self.h_out.array[0])              # already divided by the global N
        return (loss, self.h_amax.array.copy()) if want_argmax else loss

    @property
    def h2d_bytes(self):
        return self.h_in.nbytes + self.h_lab.nbytes

    @property
    def d2h_bytes(self):
        """Bytes step() reads back per call: the loss, and (one-graph path) the arg-max of every row."""
        one_graph = (self.use_graph and os.environ.get("NTB200_E2E_GRAPH", "1") != "0" and (self.dp is None or self.dp.world == 1))
        return 4 + (self.h_amax.nbytes if one_graph else 0)

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
        self.sync_replicas()

    def close(self):
        """Destroy the captured graphs (required before DataParallel.shutdown: they hold NCCL kernels)."""
        lib.nt_sync()
        if self.graph is not None:
            lib.nt_graph_destroy(self.graph)
            self.graph = None
        for g in getattr(self, "_e2e_graphs", {}).values():
            lib.nt_graph_destroy(g)
        self._e2e_graphs, self._e2e_buffers = {}, []
        self._graph_buffers = None
        if self.dp is not None and getattr(self.dp, "p2p_ready", False) and self.dp.p2p_base == self.arena.grads.ptr:
            self.dp.close_p2p()                     # the peers' mappings of this arena end with it (collective)

    def __del__(self):
        try:
            if self.graph is not None:
                lib.nt_graph_destroy(self.graph)
        except Exception:
            pass
