"""Data parallelism on real GPUs (needs >= 2 visible devices; skipped otherwise): tools/dp_check.py under torchrun.
Ranks seed their models DIFFERENTLY; after TrainStep's parameter broadcast and four steps with the gradient all-reduce
(NVLink peer-memory kernel for these small arenas, NCCL otherwise) every rank's parameters must equal the oracle's
single-process step on the full batch to 1e-4."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


@pytest.mark.parametrize("p2p", ["1", "0"])
def test_two_rank_training_matches_the_oracle(p2p):
    if _gpus() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ, NTB200_P2P=p2p)
    cp = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                         "127.0.0.1", "--master-port", "29541" if p2p == "1" else "29542", os.path.join(ROOT, "tools", "dp_check.py")],
                        cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    assert cp.returncode == 0, (cp.stdout + cp.stderr)[-3000:]
    assert cp.stdout.count("DP_CHECK") == 2 and "FAIL" not in cp.stdout
    assert ("NVLink peer-memory kernel" in cp.stdout) == (p2p == "1")
