"""Phase programs vs the per-op launches they replace (development aid; numbers quoted in DESIGN.md 4.7):
the gpt_character training step (BASELINE.json configs[2]) and the KV-cache decode step of the poetry3000-size GPT."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.generate import KVGenerator

nt.init(0)
nt.install()


def ev():
    e = ctypes.c_void_p()
    lib.nt_event_create(ctypes.byref(e))
    return e


def time_step(fused, n=300):
    B, T, D, H, L, V = 2, 5, 192, 3, 1, 26
    layers = trainer.build_gpt_layers(B, T, D, H, L, V, adam=True, seed=0)
    ts = trainer.TrainStep(layers, "gpt", (B, T), 3e-3, adam=True)
    ts._fuse = bool(fused)
    rs = np.random.RandomState(0)
    ts.load(rs.randint(0, V, (B, T)), rs.randint(0, V, (B, T)))
    for _ in range(5):
        ts.run_device()
    e0, e1 = ev(), ev()
    best = 1e9
    for _ in range(5):
        lib.nt_sync()
        lib.nt_event_record(e0)
        for _ in range(n):
            ts.run_device()
        lib.nt_event_record(e1)
        ms = ctypes.c_float(0)
        lib.nt_event_elapsed_ms(e0, e1, ctypes.byref(ms))
        best = min(best, ms.value / n)
    loss = float(ts.d_loss.numpy()[0])
    k = ts.graph_kernels
    ts.close()
    return best, k, loss


for fused in (False, True):
    ms, k, loss = time_step(fused)
    print("gpt_char step  %-9s %.4f ms/step  %2d kernels per step  (loss after the run %.5f)" % ("program" if fused else "per-op", ms, k, loss), flush=True)

for fused in (() if "--step-only" in sys.argv else ("0", "1")):
    os.environ["NTB200_FUSE"] = fused
    layers = trainer.build_gpt_layers(1, 256, 384, 6, 5, 2350, adam=False, seed=0)
    gen = KVGenerator(layers)
    prompt = np.random.RandomState(0).randint(0, 2350, (1, 30))
    gen.generate(prompt, 8)
    nt.synchronize()
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        out = gen.generate(prompt, 201)
        best = min(best, time.perf_counter() - t0)
    print("decode (D=384, 5 blocks, V=2350, B=1) %-9s 200 tokens (+prefill) in %.4f s -> %.0f tokens/s (%.4f ms/token), %d kernels per step, tail %s"
          % ("program" if fused == "1" else "per-op", best, 200 / best, best / 200 * 1e3, gen.graph_kernels, out[0, -4:].tolist()), flush=True)
    gen.close()

a, b, c, d = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_int(0)
lib.nt_fuse_stats(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(d))
print("programs recorded %d, phases %d, cluster barriers %d, CTAs per cluster %d" % (a.value, b.value, c.value, d.value))
