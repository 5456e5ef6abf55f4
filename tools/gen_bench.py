"""Generation throughput of the poetry3000-size GPT (D=384, 6 heads, 5 blocks, T=256, V=2350): KV-cache step vs the
reference's full-window recomputation per token (development aid; numbers quoted in DESIGN.md)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200.generate import KVGenerator

nt.init(0)
layers = trainer.build_gpt_layers(1, 256, 384, 6, 5, 2350, adam=False, seed=0)
gen = KVGenerator(layers)
prompt = np.random.RandomState(0).randint(0, 2350, (1, 30))
gen.generate(prompt, 8)            # eager step + graph capture outside the timed loops
nt.synchronize()
t0 = time.perf_counter()
out = gen.generate(prompt, 201)
dt = time.perf_counter() - t0
print("greedy-graph 200 tokens (+prefill) in %.3f s -> %.0f tokens/s (%.3f ms/token), %d kernels per step" % (
    dt, 200 / dt, dt / 200 * 1e3, gen.graph_kernels), flush=True)
for mode in ("kv", "full"):
    logits = gen.prefill(prompt)
    seq = prompt.copy()
    nt.synchronize()
    t0 = time.perf_counter()
    n = 200
    for _ in range(n):
        nxt = np.argmax(logits, -1)
        seq = np.concatenate([seq, nxt[:, None]], 1)
        logits = gen.step(nxt) if mode == "kv" else gen.full_logits(seq)
    dt = time.perf_counter() - t0
    print("%-4s %d tokens in %.3f s -> %.0f tokens/s (%.2f ms/token)" % (mode, n, dt, n / dt, dt / n * 1e3), flush=True)
