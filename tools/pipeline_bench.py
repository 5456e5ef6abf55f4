"""One epoch of the README ViT over an MNIST-sized synthetic set (60000 x 28 x 28 uint8), wall clock INCLUDING batch
construction: the reference's host procedure (fancy-index shuffle, slice, one-hot, fp64 -> upload) feeding TrainStep.step,
vs the resident data set of numpy_transformer_b200/data.py (one gather kernel per step).  Development aid; numbers quoted
in DESIGN.md 6.1."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200.data import ImageBatches

nt.init(0)
B, n = 100, 60000
rs = np.random.RandomState(0)
raw, lab = rs.randint(0, 256, (n, 28, 28)).astype(np.uint8), rs.randint(0, 10, n)
layers = trainer.build_vit_layers(B, seed=0)
ts = trainer.TrainStep(layers, "vit", (B, 1, 28, 28), lr=1e-3)
for _ in range(5):
    ts.step(raw[:B, None] / 255, lab[:B])

# the reference's host pipeline (transformer_of_image.py:99-139), feeding the same TrainStep
t0 = time.perf_counter()
datas, labels = raw / 255, lab.copy()
onehot = np.zeros((n, 10)); onehot[range(n), labels] = 1
t_prep = time.perf_counter() - t0
t0 = time.perf_counter()
k = np.arange(n); np.random.shuffle(k)
datas, onehot, labels = datas[k], onehot[k], labels[k]
tot = 0.0
for j in range(n // B):
    images = datas[j * B:(j + 1) * B][:, np.newaxis]
    tot += ts.step(images, labels[j * B:(j + 1) * B])
nt.synchronize()
t_host = time.perf_counter() - t0
print("host pipeline  : %.3f s/epoch -> %.0f images/s (+ %.2f s one-off /255 and one-hot), mean loss %.4f" % (
    t_host, n / t_host, t_prep, tot / (n // B)), flush=True)

loader = ImageBatches(raw, lab, B)
t0 = time.perf_counter()
loader.shuffle()
for j in range(len(loader)):
    loader.feed(j, ts)
    ts.run_device()
last = float(ts.loss)           # one host read at the end of the epoch
t_dev = time.perf_counter() - t0
print("device pipeline: %.3f s/epoch -> %.0f images/s, last loss %.4f" % (t_dev, n / t_dev, last), flush=True)
