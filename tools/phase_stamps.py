"""Where a phase program spends its time (development aid): per phase the SM-clock cycles of the body and of
the barrier behind it, from the stamps `nt_fuse_debug_stamps` makes the kernel write.  Runs the gpt_character training
step and one decode step of the poetry3000-size GPT; NTB200_FUSE_CLUSTER picks the cluster size."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["NTB200_FUSE"] = "1"
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import DeviceArray
from numpy_transformer_b200.generate import KVGenerator

OPS = ["nop", "linear_fwd", "linear_bwd_input", "linear_bwd_params", "ln_fwd", "ln_bwd", "attn_fwd", "attn_bwd", "ce_rows",
       "ce_reduce", "embed_fwd", "embed_bwd", "decode_embed", "attn_decode", "argmax_next", "increment"]
nt.init(0)
nt.install()


def cluster():
    a, b, c, d = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_int(0)
    lib.nt_fuse_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
    return d.value, b.value


def report(title, buf, n_phases, C):
    s = buf.numpy().reshape(64, 16, 3)[:n_phases, :C].astype(np.int64)
    body = s[:, :, 1] - s[:, :, 0]
    bar = s[:, :, 2] - s[:, :, 1]
    total = s[:, 0, 2] - s[:, 0, 0]
    print("%s: %d phases on %d CTAs; CTA 0 start -> end %d cycles" % (title, n_phases, C, s[-1, 0, 2] - s[0, 0, 0]))
    print("  phase  body cycles (min / max over CTAs)   barrier cycles (min / max)   CTA 0 total")
    for i in range(n_phases):
        print("  %3d   %8d / %8d              %8d / %8d          %8d" % (i, body[i].min(), body[i].max(), bar[i].min(), bar[i].max(), total[i]))
    print("  sum of max body %d, sum of min barrier %d, sum of CTA-0 totals %d cycles" % (body.max(1).sum(), bar.min(1).sum(), total.sum()), flush=True)


stamps = DeviceArray.zeros((64 * 16 * 3 * 2,), dtype=np.int32)          # int64 [64][16][3]


class I64View:
    def __init__(self, d):
        self.d = d

    def numpy(self):
        return self.d.numpy().view(np.int64)


B, T, D, H, L, V = 2, 5, 192, 3, 1, 26
layers = trainer.build_gpt_layers(B, T, D, H, L, V, adam=True, seed=0)
ts = trainer.TrainStep(layers, "gpt", (B, T), 3e-3, adam=True, use_graph=False)
rs = np.random.RandomState(0)
ts.load(rs.randint(0, V, (B, T)), rs.randint(0, V, (B, T)))
for _ in range(3):
    ts.run_device()
_, p0 = cluster()
lib.nt_fuse_debug_stamps(stamps.ptr)
ts.run_device()
lib.nt_sync()
lib.nt_fuse_debug_stamps(None)
C, p1 = cluster()
report("gpt_character step", I64View(stamps), p1 - p0, C)

layers = trainer.build_gpt_layers(1, 256, 384, 6, 5, 2350, adam=False, seed=0)
gen = KVGenerator(layers, use_graph=False)
logits = gen.prefill(rs.randint(0, 2350, (1, 30)))
for _ in range(3):
    logits = gen.step(np.argmax(logits, -1))
_, p0 = cluster()
lib.nt_fuse_debug_stamps(stamps.ptr)
logits = gen.step(np.argmax(logits, -1))
lib.nt_fuse_debug_stamps(None)
C, p1 = cluster()
report("decode step (D=384, 5 blocks, V=2350, B=1)", I64View(stamps), p1 - p0, C)
