"""Device time of the fused ViT block stack (forward / backward) on the README shape: tcgen05 kernels (NTB200_VIT_TC=1)
vs the mma.sync kernels (=0).  Development aid."""
import ctypes, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
nt.init(0); nt.install()
from numpy_transformer_b200 import fused_vit
from numpy_transformer_b200._lib import lib
from attention import attention_layer

B, N, L = int(os.environ.get("B", 100)), 49, 6
np.random.seed(0)
layers = [attention_layer(96, 4) for _ in range(L)]
x = nt.as_device(np.random.randn(B, N, 96))
dout = nt.as_device(np.random.randn(B, N, 96))
bwd = "--bwd" in sys.argv


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    ev = []
    for _ in range(2 * n):
        e = ctypes.c_void_p(); lib.nt_event_create(ctypes.byref(e)); ev.append(e)
    for i in range(n):
        lib.nt_flush_l2()
        lib.nt_event_record(ev[2 * i]); fn(); lib.nt_event_record(ev[2 * i + 1])
    lib.nt_sync()
    ms = ctypes.c_float(0); out = []
    for i in range(n):
        lib.nt_event_elapsed_ms(ev[2 * i], ev[2 * i + 1], ctypes.byref(ms)); out.append(ms.value)
    return np.median(out) * 1e3


for tc in ("1", "0"):
    os.environ["NTB200_VIT_TC"] = tc
    for want_P in (True, False):
        if tc == "0" and not want_P:
            continue
        f = lambda: fused_vit.forward_stack(layers, x, prepared=False, want_P=want_P)
        print("VIT_TC=%s want_P=%s forward (incl. weight prep): %.1f us" % (tc, want_P, timeit(f)))
        if bwd:
            f()
            print("VIT_TC=%s backward: %.1f us" % (tc, timeit(lambda: fused_vit.backward_stack(layers, dout))))

# phase timeline of CTA 0 (tcgen05 forward)
os.environ["NTB200_VIT_TC"] = "1"
dbg = nt.DeviceArray.zeros((2 * 16 * 16 * 2,), dtype=np.int32)
lib.nt_vit_tc_debug(dbg.ptr)
fused_vit.forward_stack(layers, x, want_P=False)
lib.nt_sync()
st = dbg.numpy().view(np.int64).reshape(2, 16, 16)
lib.nt_vit_tc_debug(None)
t0 = st[0, 0, 0]
names_e = ["start", "U ready", "qacc", "kacc", "vacc", "qkv rdy", "S ready", "P written", "O ready", "U1 ready", "f0acc0", "f0acc1",
           "G ready", "f1acc", "out stored"]
for bi in (0, 1, 5):
    print("block %d epilogue thread 0 (cycles since kernel's first stamp; delta):" % bi)
    prev = st[0, bi, 0]
    for k, nme in enumerate(names_e):
        print("   %-12s %8d  +%d" % (nme, st[0, bi, k] - t0, st[0, bi, k] - prev))
        prev = st[0, bi, k]
    print("  MMA thread:", [int(v - t0) for v in st[1, bi, :9]])

if bwd:
    lib.nt_vit_tc_debug(dbg.ptr)
    fused_vit.forward_stack(layers, x, want_P=False)
    dbg.fill_zero()
    fused_vit.backward_stack(layers, dout)
    lib.nt_sync()
    st = dbg.numpy().view(np.int64).reshape(2, 16, 16)
    lib.nt_vit_tc_debug(None)
    t0 = st[0, 0, 0]
    names_b = ["start", "D stored", "U1 stored", "dGacc0", "dGacc1", "G,dpre rdy", "dU1acc", "LN1bwd done", "dW0 done", "att planes",
               "S ready", "P written", "dS written", "dQK ready", "U stored", "dU ready"]
    for n in (0, 1, 5):
        print("backward iteration %d, token thread 0 (cycles; delta):" % n)
        prev = st[0, n, 0]
        for k, nme in enumerate(names_b):
            print("   %-12s %8d  +%d" % (nme, st[0, n, k] - t0, st[0, n, k] - prev))
            prev = st[0, n, k]
    print("LN1-backward sub-phases (iteration 1): after ld24, products, colsums, row sums, final, scratch store:",
          [int(st[1, 1, k] - st[0, 1, 6]) for k in range(6)])
