"""Per-op device times of one captured KV-cache decode step (poetry3000-size GPT, B=1): development aid."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.device import keep_allocations
from numpy_transformer_b200.generate import KVGenerator

nt.init(0)
B = int(os.environ.get("GEN_B", "1"))
layers = trainer.build_gpt_layers(B, 256, 384, 6, 5, 2350, adam=False, seed=0)
gen = KVGenerator(layers, use_graph=False)
prompt = np.random.RandomState(0).randint(0, 2350, (B, 100))
gen.prefill(prompt)
gen.step(prompt[:, 0])
lib.nt_sync()
with keep_allocations() as keep:
    lib.profile_begin(external=True)
    lib.nt_graph_begin()
    gen._enqueue_step()
    recs = lib.profile_pause()
    g, nk = ctypes.c_void_p(), ctypes.c_int(0)
    lib.nt_graph_end(ctypes.byref(g), ctypes.byref(nk))
for _ in range(5):
    lib.nt_graph_launch(g)
lib.nt_sync()
tot = {}
for name, args, ms in lib.profile_read(recs):
    key = name + " " + str(args[-6:])
    n, t = tot.get(key, (0, 0.0))
    tot[key] = (n + 1, t + ms)
total = sum(t for _, t in tot.values())
print("%d kernels, %.1f us summed over ops" % (nk.value, total * 1e3))
for k, (n, t) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%6.1f us  x%-3d %5.1f us each  %s" % (t * 1e3, n, t * 1e3 / n, k))
