import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def golden_tree(d, prefix, like):
    """Rebuild a checkpoint tree from the flat <prefix>_NNN arrays of a golden file."""
    from oracle import models as M
    n = len(M.tree_leaves(like))
    leaves = [d["%s_%03d" % (prefix, i)] for i in range(n)]
    return M.tree_unflatten(like, leaves)


def rel_err(a, b):
    """Tolerance metric used across the suite: max|a-b| / max|b| per tensor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = np.max(np.abs(b))
    if denom == 0:
        return float(np.max(np.abs(a)))
    return float(np.max(np.abs(a - b)) / denom)


@pytest.fixture(scope="session")
def prims():
    return load_golden("prims.npz")
