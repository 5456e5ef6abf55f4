"""Device time of the attention core at the scaled-ViT shape (B=1024, N=49, 6 heads x 64): tcgen05 kernels (NTB200_ATT_TC=1,
default) vs the mma.sync BF16x3 kernels (=0).  Development aid."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
nt.init(0)
from numpy_transformer_b200 import ops
from numpy_transformer_b200._lib import lib

B, N, H, dh = int(os.environ.get("B", 1024)), int(os.environ.get("N", 49)), 6, 64
rs = np.random.RandomState(0)
qkv = nt.as_device(rs.randn(B, N, 3 * H * dh).astype(np.float32))
dout = nt.as_device(rs.randn(B, N, H * dh).astype(np.float32))


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ev = []
    for _ in range(2 * n):
        e = ctypes.c_void_p(); lib.nt_event_create(ctypes.byref(e)); ev.append(e)
    for i in range(n):
        lib.nt_event_record(ev[2 * i]); fn(); lib.nt_event_record(ev[2 * i + 1])
    lib.nt_sync()
    ms = ctypes.c_float(0); out = []
    for i in range(n):
        lib.nt_event_elapsed_ms(ev[2 * i], ev[2 * i + 1], ctypes.byref(ms)); out.append(ms.value)
    return np.median(out) * 1e3


res = {}
def fwd():
    res["o"] = ops.attention_fwd(qkv, H, None, ops.MASK_NONE, split_out=True)
def bwd():
    o, P, _ = res["o"]
    ops.attention_bwd(dout, qkv, P, H, ops.MASK_NONE, split_out=True, att_out=o)
print("NTB200_ATT_TC=%s B=%d N=%d H=%d dh=%d: forward %.1f us, backward %.1f us" % (
    os.environ.get("NTB200_ATT_TC", "1"), B, N, H, dh, timeit(fwd), timeit(bwd)))
