"""TEST INFRASTRUCTURE — an unmodified copy of the reference tree under oracle/_ref/tree, plus the fixtures its
trainer scripts need, so that the scripts themselves can be executed (a) on the reference's own NumPy layers and
(b) on the B200 drop-in, and their log lines compared.

* `materialize()` is the "build" of oracle/_ref for a Python reference: it copies the reference's sources and the
  two small corpora the GPT trainers read from /root/reference (present in the build container only) into
  oracle/_ref/tree.  oracle/_ref is git-ignored (never in history) and NOT gpurun-ignored (it travels to the GPU
  box like a built .so).  Nothing is edited; tests/golden/ref_scripts/MANIFEST.json pins the sha256 of every file.
* `stage(work)` gives a script its own writable copy (they write logs / checkpoints next to themselves:
  transformer_of_image.py:20-22,52-53,196; gpt/gpt_train_potry3000.py:2-4,160,371).
* `make_mnist(root, ...)` writes synthetic idx files where torchvision's existence check looks
  (transformer_of_image.py:94-97) — there is no network for the real data set.
* `make_poetry_checkpoint(work, ...)` produces gpt/model/gpt_poetry3000_last4.pkl with the REFERENCE's own
  constructors and save_model() walk (gpt/gpt_train_potry3000.py:212-233 exits without it).
* `run_script(work, script, backend, seed)` executes the script as __main__ in a subprocess: backend "numpy" =
  the reference's layers, backend "b200" = `python -m numpy_transformer_b200.run` (the product's launcher).

Only tests/, __graft_entry__ and bench.py's CPU legs may import this module.
"""
import gzip  # noqa: F401  (kept for symmetry with torchvision's reader; the raw files are uncompressed)
import hashlib
import json
import os
import shutil
import struct
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REFERENCE = "/root/reference"
TREE = os.path.join(HERE, "_ref", "tree")
MPL_STUB = os.path.join(HERE, "_mpl_stub")
MANIFEST = os.path.join(ROOT, "tests", "golden", "ref_scripts", "MANIFEST.json")

# what travels: every Python source except the 500 KB scratch file kkk.py, and the corpora of the two GPT trainers
_DATA = ("dataset/train_3000.txt", "dataset/train_3000.json", "dataset/alphabet.json", "dataset/alphabet.txt")


def _wanted(src_root):
    out = []
    for d, _, files in os.walk(src_root):
        if ".git" in d.split(os.sep):
            continue
        for f in files:
            rel = os.path.relpath(os.path.join(d, f), src_root)
            if rel.endswith(".py") and rel != "kkk.py":
                out.append(rel)
    out += [p for p in _DATA if os.path.exists(os.path.join(src_root, p))]
    return sorted(out)


def _sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        h.update(f.read())
    return h.hexdigest()


def materialize(src=REFERENCE, dst=TREE):
    """Copy the reference (unmodified) into oracle/_ref/tree; returns {relpath: sha256}."""
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    manifest = {}
    for rel in _wanted(src):
        t = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(t), exist_ok=True)
        shutil.copyfile(os.path.join(src, rel), t)
        manifest[rel] = _sha(t)
    return manifest


def available():
    return os.path.exists(os.path.join(TREE, "transformer_of_image.py"))


def tree_manifest(tree=TREE):
    return {rel: _sha(os.path.join(tree, rel)) for rel in _wanted(tree)}


def stage(work):
    """A private writable copy of the tree for one script run."""
    dst = os.path.join(work, "tree")
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(TREE, dst)
    return dst


def _write_idx(path, arr):
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    with open(path, "wb") as f:
        f.write(struct.pack(">BBBB", 0, 0, 8, arr.ndim))
        for s in arr.shape:
            f.write(struct.pack(">I", s))
        f.write(arr.tobytes())


def make_mnist(tree, n_train=200, n_test=1000, seed=7):
    """Synthetic MNIST idx files under <tree>/dataset/MNIST/raw (blurred class-dependent blobs + noise, so the
    loss actually moves).  torchvision.datasets.MNIST(root, download=True) finds them and skips the download."""
    raw = os.path.join(tree, "dataset", "MNIST", "raw")
    os.makedirs(raw, exist_ok=True)
    rs = np.random.RandomState(seed)
    yy, xx = np.mgrid[0:28, 0:28]
    for split, n in (("train", n_train), ("t10k", n_test)):
        lab = rs.randint(0, 10, n).astype(np.uint8)
        cx = 6 + 16 * ((lab % 5) / 4.0)
        cy = 8 + 12 * (lab // 5)
        img = 200.0 * np.exp(-((xx[None] - cx[:, None, None]) ** 2 + (yy[None] - cy[:, None, None]) ** 2) / 18.0)
        img = np.clip(img + rs.randint(0, 56, img.shape), 0, 255).astype(np.uint8)
        _write_idx(os.path.join(raw, "%s-images-idx3-ubyte" % split), img)
        _write_idx(os.path.join(raw, "%s-labels-idx1-ubyte" % split), lab)
    return raw


def _env(extra_path=()):
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([ROOT] + list(extra_path) + [env.get("PYTHONPATH", "")]).rstrip(os.pathsep)
    env.setdefault("OMP_NUM_THREADS", str(os.cpu_count() or 1))
    return env


_CKPT_SNIPPET = r"""
import sys, pickle, numpy as np
tree, out, seed, jk, nblocks = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
sys.path.insert(0, sys.argv[6]); sys.path.insert(0, tree + '/gpt'); sys.path.insert(0, tree)
from PatchEmbed import Position_Embedding
from attdecoderblock import attdecoderblock_layer
from net.layernorm import layer_norm
from gpt.classify_gpt import classify_layer
np.random.seed(seed)
D, T, V, B = 384, 256, 2350, 64
layers = [Position_Embedding(T, V, D, adam=True, float32=True, float16=False)]
layers += [attdecoderblock_layer(D, 6, adam=True, float32=True, float16=False) for _ in range(nblocks)]
layers += [layer_norm(D, adam=True, float32=True, float16=False),
           classify_layer(D, B, 1, V, cls_token=False, adam=True, relu=False, float32=True, float16=False)]
allmodel = [l.save_model() for l in layers if 'restore_model' in dir(l) and 'save_model' in dir(l)]
allmodel.append(jk)
with open(out, 'wb') as f:
    pickle.dump(allmodel, f)
"""


def make_poetry_checkpoint(tree, seed=11, jk=6598, nblocks=5):
    """gpt/model/gpt_poetry3000_last4.pkl from the reference's own constructors + save_model() (CPU, seconds).
    jk = the step counter the trainer resumes from (it stops after step 6601: gpt_train_potry3000.py:237-241)."""
    out = os.path.join(tree, "gpt", "model", "gpt_poetry3000_last4.pkl")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call([sys.executable, "-c", _CKPT_SNIPPET, tree, out, str(seed), str(jk), str(nblocks), MPL_STUB],
                          env=_env())
    return out


_NUMPY_RUNNER = r"""
import sys, runpy, numpy as np
tree, script, seed, stub = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
sys.path.insert(0, stub)
if not hasattr(np.lib, 'pad'):
    np.lib.pad = np.pad          # NumPy 2 dropped the alias net/Convolution.py:114 uses
np.random.seed(seed)
sys.argv = [script]
runpy.run_path(script, run_name='__main__')
"""


def run_script(tree, script, backend, seed=0, timeout=3600):
    """Execute <tree>/<script> as __main__ (cwd = tree).  Returns the CompletedProcess."""
    path = os.path.join(tree, script)
    if backend == "numpy":
        # `python script.py` semantics: the script's directory leads sys.path, the scripts add the tree root themselves
        extra = [os.path.dirname(path), tree, os.path.join(tree, "gpt")]
        cmd = [sys.executable, "-c", _NUMPY_RUNNER, tree, path, str(seed), MPL_STUB]
    elif backend == "b200":
        extra = [MPL_STUB]
        cmd = [sys.executable, "-m", "numpy_transformer_b200.run", "--seed", str(seed), path]
    else:
        raise ValueError(backend)
    return subprocess.run(cmd, cwd=tree, env=_env(extra), timeout=timeout, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                          text=True)


# ---------------------------------------------------------------- the three trainer runs (shared by the golden
# generator and the GPU test, so both arms see identical fixtures)
RUNS = {
    # name: (script, log file the script writes, seed)
    "vit": ("transformer_of_image.py", "log/log_transformer_of_image__pf_pn_fixed.txt", 3),
    "gpt_char": ("gpt_character/gpt_train_english_char.py", "gpt_character/log/log_gpt_alphabet.txt", 5),
    "gpt_poetry": ("gpt/gpt_train_potry3000.py", "gpt/log/log_gpt_poetry3000.txt", 9),
}


def prepare(name, work):
    tree = stage(work)
    if name == "vit":
        make_mnist(tree)
    elif name == "gpt_poetry":
        make_poetry_checkpoint(tree)
        log = os.path.join(tree, RUNS[name][1])
        if os.path.exists(log):
            os.remove(log)               # the trainer appends ('a+'); the reference ships its own historical log
    return tree


def run(name, work, backend, timeout=3600):
    script, log, seed = RUNS[name]
    tree = prepare(name, work)
    cp = run_script(tree, script, backend, seed=seed, timeout=timeout)
    logpath = os.path.join(tree, log)
    text = open(logpath, encoding="utf-8").read() if os.path.exists(logpath) else ""
    return cp, text, tree


def parse_log(text):
    """-> list of dicts of the numeric fields of every 'epoch:..., loss: ...' line (plus the sampled text lines)."""
    import re
    rows, extra = [], []
    for line in text.splitlines():
        m = re.match(r"epoch:\s*(\d+), lr: ([\d.eE+-]+), loss: ([\d.eE+-]+|nan|inf), iters: ([\w_]+), precision: ([\d.]+)"
                     r"(?:, valpre: ([\d.]+))?", line)
        if m:
            rows.append(dict(epoch=int(m.group(1)), lr=float(m.group(2)), loss=float(m.group(3)), iters=m.group(4),
                             precision=float(m.group(5)), valpre=float(m.group(6)) if m.group(6) else None))
        else:
            extra.append(line)
    return rows, extra


if __name__ == "__main__":
    # python -m oracle.ref_tree            -> (re)build oracle/_ref/tree and print the manifest
    # python -m oracle.ref_tree golden     -> also run the three trainers on the reference's NumPy layers and write
    #                                         tests/golden/ref_scripts/<name>.log (+ MANIFEST.json)
    man = materialize()
    print("oracle/_ref/tree: %d files" % len(man))
    if len(sys.argv) > 1 and sys.argv[1] == "golden":
        import tempfile
        import time
        os.makedirs(os.path.dirname(MANIFEST), exist_ok=True)
        json.dump(man, open(MANIFEST, "w"), indent=1, sort_keys=True)
        names = sys.argv[2:] or list(RUNS)
        for name in names:
            with tempfile.TemporaryDirectory() as work:
                t0 = time.time()
                cp, text, _ = run(name, work, "numpy", timeout=4 * 3600)
                print("%s: rc=%d, %.0f s, %d log lines" % (name, cp.returncode, time.time() - t0, len(text.splitlines())))
                if cp.returncode != 0:
                    print(cp.stdout[-4000:])
                    raise SystemExit(1)
                open(os.path.join(os.path.dirname(MANIFEST), name + ".log"), "w", encoding="utf-8").write(text)
