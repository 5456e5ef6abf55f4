"""Does tcgen05 kind::tf32 truncate or round its fp32 inputs?  Compare the single-pass GEMM against fp64
references built from truncated / round-to-nearest tf32 inputs (development aid)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import ops, DeviceArray


def trunc(a):
    b = a.astype(np.float32).view(np.uint32) & np.uint32(0xFFFFE000)
    return b.view(np.float32).astype(np.float64)


def rn(a):
    u = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32).astype(np.float64)


nt.init(0)
nt.set_gemm_mode(2)
rs = np.random.RandomState(0)
x, w = rs.randn(256, 64), rs.randn(64, 128)
y = np.asarray(ops.linear_fwd(DeviceArray.from_host(x), DeviceArray.from_host(w)), dtype=np.float64)
for name, f in (("trunc", trunc), ("round", rn)):
    ref = f(x) @ f(w)
    print(name, "max rel err", float(np.max(np.abs(y - ref)) / np.max(np.abs(ref))))
