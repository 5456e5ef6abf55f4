"""World-size-2 gloo tests (CPU) for the data-parallel host logic: sharding, bucket bounds,
unique-id exchange, and the global-N gradient normaliser (sum of shard gradients == full-batch
gradient), with the oracle as the checker."""
import os
import sys
import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_and_bucket_math():
    from numpy_transformer_b200.parallel import shard_range, bucket_bounds
    assert [shard_range(1024, r, 8) for r in (0, 7)] == [(0, 128), (896, 1024)]
    with pytest.raises(ValueError):
        shard_range(100, 0, 8)
    for n, b in [(847498 + 2, 4), (16, 8), (4, 3), (1000, 1)]:
        bb = bucket_bounds(n, b)
        assert bb[0][0] == 0 and bb[-1][1] == n
        assert all(lo % 4 == 0 for lo, _ in bb)
        assert all(bb[i][1] == bb[i + 1][0] for i in range(len(bb) - 1))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from numpy_transformer_b200.parallel import DataParallel, shard_range
    from oracle import models as M
    dp = DataParallel(rank, world, backend="gloo")
    uid = dp.exchange_unique_id(lambda: bytes(range(128)))
    assert uid == bytes(range(128))
    cfg = M.ViTConfig(batch=4, embed_dim=16, heads=2, blocks=1)
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    lo, hi = shard_range(cfg.batch, rank, world)
    logits, acts = M.vit_forward(params, images[lo:hi], cfg)
    # local softmax / gradient with the GLOBAL normaliser (loss.py:49-50 divides by predict.shape[0])
    z = logits - logits.max(-1, keepdims=True)
    p = np.exp(z) / np.exp(z).sum(-1, keepdims=True)
    partial = (p - onehot[lo:hi]) / cfg.batch
    grads, _ = M.vit_backward(params, acts, partial, cfg)
    flat = np.concatenate([g.reshape(-1) for g in M.tree_leaves(grads)])
    t = torch.from_numpy(flat.copy())
    dist.all_reduce(t)
    full = M.vit_step(params, images, onehot, 0.1, cfg)
    ref = np.concatenate([g.reshape(-1) for g in M.tree_leaves(full["grads"])])
    err = float(np.max(np.abs(t.numpy() - ref)) / np.max(np.abs(ref)))
    mx = dp.max_over_ranks(float(rank))
    dp.barrier()
    q.put((rank, err, mx))
    dp.shutdown()


def test_two_rank_gradient_equivalence_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err, mx in res:
        assert err < 1e-12
        assert mx == 1.0
