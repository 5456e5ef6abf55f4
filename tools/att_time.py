"""Device time of the attention core at the scaled-ViT shape (B=1024, N=49, 6 heads x 64): tcgen05 kernels (NTB200_ATT_TC=1,
default) vs the mma.sync BF16x3 kernels (=0).  Development aid."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
nt.init(0)
from numpy_transformer_b200 import ops
from numpy_transformer_b200._lib import lib

B, N, H, dh = int(os.environ.get("B", 1024)), int(os.environ.get("N", 49)), int(os.environ.get("H", 6)), 64
KIND = ops.MASK_CAUSAL if os.environ.get("CAUSAL", "0") == "1" else ops.MASK_NONE
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
PLANES = os.environ.get("PLANES", "1") == "1" and N > 64
if PLANES:
    planes = ops.split_bf16(qkv, B * N, 3 * H * dh)[:2]
def fwd():
    if PLANES:
        qkv._split = planes
    res["o"] = ops.attention_fwd(qkv, H, None, KIND, split_out=True)
if PLANES:
    dout._split = ops.split_bf16(dout, B * N, H * dh)[:2]
def bwd():
    o, P, _ = res["o"]
    ops.attention_bwd(dout, qkv, P, H, KIND, split_out=True, att_out=o)
print("PLANES=%s " % PLANES, end="")
print("NTB200_ATT_TC=%s NTB200_ATT_TCL=%s B=%d N=%d H=%d dh=%d causal=%s: forward %.1f us, backward %.1f us" % (
    os.environ.get("NTB200_ATT_TC", "0"), os.environ.get("NTB200_ATT_TCL", "1"), B, N, H, dh, KIND == ops.MASK_CAUSAL, timeit(fwd), timeit(bwd)))

if N > 64 and os.environ.get("NTB200_ATT_TCL", "1") != "0":
    # phase timeline of CTA 0 of the tcgen05 forward (csrc/attention_tcl.cu)
    dbg = nt.DeviceArray.zeros((8 * 2 * 16 * 2,), dtype=np.int32)
    lib.nt_attention_tcl_debug(dbg.ptr)
    fwd()
    lib.nt_sync()
    st = dbg.numpy().view(np.int64).reshape(8, 2, 16)
    lib.nt_attention_tcl_debug(None)
    names = ["start", "q,k stored", "v stored", "S issued", "S in regs", "max done", "sum done", "P stored", "O issued", "O in regs",
             "out stored", "P to global", "q,k fetched", "cta barrier", "O complete"]
    t0 = st[0, 0, 0]
    for n in range(int(os.environ.get('ITEMS', 3))):
        for th in (0, 1):
            if st[n, th, 0] == 0:
                continue
            print("item %d of CTA 0, thread %s (cycles since start; delta):" % (n, "0" if th == 0 else "480"))
            prev = st[n, th, 0]
            print("   " + "  ".join("%s %d(+%d)" % (nm, st[n, th, k] - t0, st[n, th, k] - (st[n, th, k - 1] if k else prev))
                                    for k, nm in enumerate(names)))

    if PLANES and os.environ.get("NTB200_ATT_TCL_BWD", "1") != "0":
        dbg.fill_zero()
        lib.nt_attention_tcl_bwd_debug(dbg.ptr)
        bwd()
        lib.nt_sync()
        st = dbg.numpy().view(np.int64).reshape(8, 2, 16)
        lib.nt_attention_tcl_bwd_debug(None)
        names = ["pair start", "loads issued, delta", "S dP issued", "S dP in regs", "P stored", "dV done", "dS stored", "dQ dK done",
                 "dK dV stored", "dQ stored"]
        t0 = st[0, 0, 0]
        for n in range(int(os.environ.get("ITEMS", 3)) * 2):
            for th in (0, 1):
                if st[n, th, 0] == 0:
                    continue
                print("backward pair %d of CTA 0, thread %s:" % (n, "0" if th == 0 else "480"))
                print("   " + "  ".join("%s %d(+%d)" % (nm, st[n, th, k] - t0, st[n, th, k] - st[n, th, k - 1] if k else 0)
                                        for k, nm in enumerate(names) if st[n, th, k]))
