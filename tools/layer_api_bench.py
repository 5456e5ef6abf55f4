"""The reference's own training loop (transformer_of_image.py:141-159: per-layer forward, cross_entropy_loss, argmax,
per-layer backward / update / setzero, all from Python, eager launches) on the drop-in classes, README ViT shape.
This is what an UNCHANGED trainer script gets; TrainStep (bench.py) is the captured-graph fast path.  Development aid."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
nt.install()
nt.init(0)
from numpy_transformer_b200 import trainer
from net.loss import cross_entropy_loss

B = 100
layers = trainer.build_vit_layers(B, seed=0)
rs = np.random.RandomState(1)
images = rs.rand(B, 1, 28, 28)
lab = rs.randint(0, 10, B)
onehot = np.zeros((B, 10)); onehot[np.arange(B), lab] = 1


def step():
    x = images
    for l in layers:
        x = l.forward(x)
    loss, delta, predict = cross_entropy_loss(x, onehot)
    p = np.argmax(predict, axis=-1)
    precision = np.sum(lab == p) / len(lab)
    for l in reversed(layers):
        delta = l.backward(delta)
        l.update(1e-3)
        l.setzero()
    return float(loss), precision


for _ in range(5):
    step()
nt.synchronize()
t0 = time.perf_counter()
n = 50
for _ in range(n):
    loss, prec = step()
nt.synchronize()
dt = (time.perf_counter() - t0) / n
print("layer-API loop: %.3f ms/step -> %.0f images/s (loss %.4f, launches/step %d)" % (
    dt * 1e3, B / dt, loss, 0), flush=True)
c0 = nt.launch_count(); step(); print("kernel launches per step:", nt.launch_count() - c0)
