"""Per-layer parameter arenas for the composite layers of the drop-in mirror.

The reference's trainers call `update(lr)` and `setzero()` on every layer object after its backward
(transformer_of_image.py:156-159).  A ViT block owns ten parameter tensors, so the literal translation is ten update
kernels and ten memsets per block and step — 60 % of the launches of the unchanged per-layer loop.  The first `update` /
`setzero` of a composite layer re-homes its tensors into one flat arena (trainer.ParamArena; the attributes become views,
exactly what TrainStep does for the whole model), after which update is ONE kernel and setzero ONE memset.  If a TrainStep
later moves the tensors into its own arena the layer notices (the views no longer live in its buffer) and falls back to
the per-tensor path."""
from ._lib import lib


class LayerArenaMixin:
    _la = None

    def _layer_arena(self):
        a = self._la
        if a is None:
            from .trainer import ParamArena
            if not self._slots():
                return None
            a = self._la = ParamArena([self], bool(getattr(self, "adam", False)))
        obj, pn, gn = a.slots[0][:3]
        if getattr(obj, pn)._buf is not a.params._buf or getattr(obj, gn)._buf is not a.grads._buf:
            return None                                   # re-homed elsewhere (TrainStep): per-tensor path
        return a

    def _arena_update(self, lr):
        """-> True when the whole layer was updated by one kernel (SGD: fullconnect.py:150-153; Adam*: :134-149)."""
        a = self._layer_arena()
        if a is None:
            return False
        owners = list({id(s[0]): s[0] for s in a.slots}.values())
        if a.m is not None:
            ts = {int(o.t) for o in owners}
            if len(ts) != 1:
                return False                              # children stepped separately: keep their own counters
            lib.nt_adam_star(a.params.ptr, a.grads.ptr, a.m.ptr, a.v.ptr, a.n, float(lr), None, ts.pop(), None, 0)
            for o in owners:
                o.t += 1
        else:
            lib.nt_sgd(a.params.ptr, a.grads.ptr, a.n, float(lr), None, 0)
        for o in owners:
            if hasattr(o, "_wsplit"):
                o._wsplit = None
        return True

    def _arena_setzero(self):
        a = self._layer_arena()
        if a is None:
            return False
        lib.nt_memset(a.grads.ptr, 0, 4 * a.n)
        return True
