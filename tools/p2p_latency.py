"""Latency of the NVLink peer-memory all-reduce kernel vs the NCCL call for small gradient ranges (run under torchrun)."""
import ctypes, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200._lib import lib
from numpy_transformer_b200.parallel import DataParallel

dp = DataParallel()
nt.init(int(os.environ.get("LOCAL_RANK", "0")))
dp.init_nccl()
n = 1 << 20
buf = nt.DeviceArray.zeros((n,))
assert dp.init_p2p(buf)
ev = [ctypes.c_void_p() for _ in range(2)]
for e in ev:
    lib.nt_event_create(ctypes.byref(e))
ms = ctypes.c_float(0)
for size in (4, 1024, 16384, 131072, 425984, 1 << 20):
    for name, fn in (("p2p ", lambda: lib.nt_p2p_allreduce_sum(0, size, 0)),
                     ("nccl", lambda: (lib.nt_dp_allreduce_sum(buf.ptr, size), lib.nt_dp_join()))):
        for _ in range(10):
            fn()
        lib.nt_sync(); dp.barrier()
        lib.nt_event_record(ev[0])
        for _ in range(200):
            fn()
        lib.nt_event_record(ev[1])
        lib.nt_sync()
        lib.nt_event_elapsed_ms(ev[0], ev[1], ctypes.byref(ms))
        t = dp.max_over_ranks(ms.value / 200 * 1e3)
        if dp.rank == 0:
            print("world %d  %s %8d floats (%7.1f KB): %7.2f us per call" % (dp.world, name, size, size * 4 / 1024, t), flush=True)
# correctness: ranks hold rank+1 everywhere -> sum = world (world + 1) / 2
buf.set(np.full((n,), dp.rank + 1.0, dtype=np.float32))
lib.nt_sync(); dp.barrier()
lib.nt_p2p_allreduce_sum(0, n, 0)
lib.nt_sync()
got = buf.numpy()
want = dp.world * (dp.world + 1) / 2
assert np.all(got == want), (got[:4], want)
if dp.rank == 0:
    print("p2p all-reduce of %d floats: exact" % n)
dp.barrier()
dp.shutdown()
