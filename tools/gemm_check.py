"""On-GPU check of the tcgen05 GEMM paths against fp64 NumPy (development aid, not a test).
usage: python tools/gemm_check.py <which: nt|nn|tn|all> <mode: 1|2> [M K N]"""
import sys
import os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import ops, DeviceArray


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - b)) / (np.max(np.abs(b)) + 1e-30))


def main():
    which, mode = sys.argv[1], int(sys.argv[2])
    shapes = [tuple(int(x) for x in sys.argv[3:6])] if len(sys.argv) >= 6 else [
        (128, 32, 128), (128, 64, 128), (256, 96, 96), (4900, 96, 288), (4900, 192, 96), (4900, 16, 96), (100, 4704, 96),
        (1000, 384, 1152), (77, 40, 200)]
    nt.init(0)
    for (M, K, N) in shapes:
        rs = np.random.RandomState(M + K + N)
        x = rs.randn(M, K)
        w = rs.randn(K, N) / np.sqrt(K)
        b = rs.randn(N)
        dy = rs.randn(M, N)
        dx_, dw_ = DeviceArray.from_host(x), DeviceArray.from_host(w)
        db_, ddy = DeviceArray.from_host(b), DeviceArray.from_host(dy)
        x32, w32, dy32 = x.astype(np.float32).astype(np.float64), w.astype(np.float32).astype(np.float64), dy.astype(np.float32).astype(np.float64)
        out = {}
        nt.set_gemm_mode(mode)
        if which in ("nt", "all"):
            r = ops.linear_bwd_input(ddy, dw_)
            out["nt(dX)"] = rel(r, dy32 @ w32.T)
        if which in ("nn", "all"):
            r = ops.linear_fwd(dx_, dw_, db_)
            out["nn(fwd)"] = rel(r, x32 @ w32 + b)
        if which in ("tn", "all"):
            g = DeviceArray.zeros((K, N))
            gb = DeviceArray.zeros((N,))
            ops.linear_bwd_params(dx_, ddy, g, gb)
            out["tn(dW)"] = rel(g, x32.T @ dy32)
            out["db"] = rel(gb, dy32.sum(0))
        nt.synchronize()
        print("mode", mode, (M, K, N), " ".join("%s=%.2e" % kv for kv in out.items()), flush=True)


main()
