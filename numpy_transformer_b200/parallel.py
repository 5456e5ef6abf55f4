"""Data parallelism: one process per GPU, batch sharded on axis 0, full replicas of
parameters / Adam* state, one bucketed NCCL all-reduce (SUM) of the flat gradient arena per step.
torch.distributed is used only for rendezvous (barrier, unique-id broadcast, max-over-ranks timing);
the collective itself is issued by libntb200 on its own comm stream so it can live inside the step
graph and overlap the tail of backward."""
import os
import numpy as np


def shard_range(global_batch, rank, world):
    """Contiguous equal shards of the global batch (axis 0).  Requires divisibility so every rank
    runs the same captured graph."""
    if global_batch % world:
        raise ValueError("global batch %d is not divisible by world size %d" % (global_batch, world))
    per = global_batch // world
    return rank * per, (rank + 1) * per


def bucket_bounds(n, buckets, align=4):
    """Split [0,n) into `buckets` contiguous ranges whose boundaries are multiples of `align` floats."""
    buckets = max(1, min(int(buckets), max(1, n // align)))
    per = (n // buckets + align - 1) // align * align
    out, lo = [], 0
    while lo < n:
        hi = min(n, lo + per)
        out.append((lo, hi))
        lo = hi
    return out


class DataParallel:
    def __init__(self, rank=None, world=None, backend=None):
        import torch.distributed as dist
        self.dist = dist
        self.rank = int(os.environ.get("RANK", 0)) if rank is None else rank
        self.world = int(os.environ.get("WORLD_SIZE", 1)) if world is None else world
        self.nccl_ready = False
        if self.world > 1 and not dist.is_initialized():
            import torch
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", self.rank)))
            dist.init_process_group(backend=backend, rank=self.rank, world_size=self.world)

    def exchange_unique_id(self, make_id):
        """rank 0 creates the 128-byte NCCL id, everyone receives it (object broadcast)."""
        obj = [make_id() if self.rank == 0 else None]
        if self.world > 1:
            self.dist.broadcast_object_list(obj, src=0)
        return obj[0]

    def init_nccl(self):
        """Create libntb200's own communicator (after device.init)."""
        import ctypes
        from ._lib import lib
        if self.world == 1 or self.nccl_ready:
            return

        def make_id():
            buf = ctypes.create_string_buffer(128)
            lib.nt_dp_unique_id(buf)
            return buf.raw
        uid = self.exchange_unique_id(make_id)
        lib.nt_dp_init(ctypes.create_string_buffer(uid, 128), self.rank, self.world)
        self.nccl_ready = True

    def init_p2p(self, grads):
        """Map every rank's gradient arena into this process (CUDA IPC over NVLink) for the small-message all-reduce
        (nt_p2p_allreduce_sum).  `grads` must be a whole device allocation (the flat arena is).  Collective."""
        import ctypes
        import os
        from ._lib import lib
        max_floats = int(float(os.environ.get("NTB200_P2P_MAX_MB", "16")) * (1 << 20) / 4)
        if (self.world == 1 or self.world > 8 or os.environ.get("NTB200_P2P", "1") == "0" or getattr(self, "p2p_ready", False)
                or grads.size > 4 * max_floats):
            return False
        flags = ctypes.c_void_p()
        lib.nt_p2p_flags_alloc(ctypes.byref(flags))
        hg, hf = ctypes.create_string_buffer(64), ctypes.create_string_buffer(64)
        lib.nt_p2p_ipc_handle(grads.ptr, hg)
        lib.nt_p2p_ipc_handle(flags, hf)
        gathered = [None] * self.world
        self.dist.all_gather_object(gathered, (self.rank, hg.raw, hf.raw, grads.size))
        gathered.sort(key=lambda t: t[0])
        assert all(t[3] == grads.size for t in gathered), "gradient arenas differ in size across ranks"
        allg = ctypes.create_string_buffer(b"".join(t[1] for t in gathered), 64 * self.world)
        allf = ctypes.create_string_buffer(b"".join(t[2] for t in gathered), 64 * self.world)
        lib.nt_p2p_setup(self.rank, self.world, grads.ptr, allg, flags, allf)
        self.p2p_ready, self.p2p_base, self.p2p_size = True, grads.ptr, grads.size
        self.p2p_max_floats = int(float(os.environ.get("NTB200_P2P_MAX_MB", "16")) * (1 << 20) / 4)
        self.barrier()
        return True

    def close_p2p(self):
        from ._lib import lib
        if getattr(self, "p2p_ready", False):
            self.barrier()
            lib.nt_p2p_shutdown()
            self.p2p_ready = False
            self.barrier()

    def allreduce_range(self, flat, lo, hi, side=False):
        """Sum grads[lo:hi) over the ranks, in place.  Small ranges of the IPC-mapped arena go through the one-kernel NVLink
        reduction, everything else through NCCL.  side=True: on the comm stream, to be joined later (nt_dp_join)."""
        from ._lib import lib
        if (getattr(self, "p2p_ready", False) and flat.ptr == self.p2p_base and hi - lo <= self.p2p_max_floats
                and lo % 4 == 0 and (hi - lo) % 4 == 0):
            lib.nt_p2p_allreduce_sum(lo, hi - lo, 1 if side else 0)
            return "p2p"
        lib.nt_dp_allreduce_sum(flat.ptr + 4 * lo, hi - lo)
        return "nccl"

    def allreduce_buckets(self, flat, buckets=1):
        for lo, hi in bucket_bounds(flat.size, buckets):
            self.allreduce_range(flat, lo, hi, side=True)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()

    def max_over_ranks(self, value):
        if self.world == 1:
            return float(value)
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(self, value):
        if self.world == 1:
            return float(value)
        import torch
        dev = "cuda" if self.dist.get_backend() == "nccl" else "cpu"
        t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def shutdown(self):
        from ._lib import lib
        if getattr(self, "p2p_ready", False):
            lib.nt_p2p_shutdown()
            self.p2p_ready = False
        if self.nccl_ready:
            lib.nt_dp_shutdown()
            self.nccl_ready = False
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()
