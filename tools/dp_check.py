"""Data-parallel parity on real GPUs (run under torchrun, one rank per GPU):
each rank takes its shard of ONE global batch, gradients are all-reduced over NCCL by libntb200,
and the post-update parameters of every rank must match the oracle's single-process step on the
full batch (SURVEY.md §4 'DP test').  Usage:
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/dp_check.py
"""
import os
import sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy_transformer_b200 as nt
from numpy_transformer_b200 import trainer
from numpy_transformer_b200.parallel import DataParallel, shard_range
from oracle import models as M


def main():
    dp = DataParallel()
    nt.init(int(os.environ.get("LOCAL_RANK", "0")))
    dp.init_nccl()
    world, rank = dp.world, dp.rank
    worst_all = 0.0
    for dim in (48, 64):            # 48: per-op path; 64: fused ViT-block kernels (csrc/vit_block.cu)
        worst_all = max(worst_all, run(dp, world, rank, dim))
    dp.shutdown()
    sys.exit(0 if worst_all < 1e-4 else 1)


def run(dp, world, rank, dim):
    per = 6
    cfg = M.ViTConfig(batch=per * world, embed_dim=dim, heads=4, blocks=2)
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    lo, hi = shard_range(cfg.batch, rank, world)
    # every rank seeds differently: TrainStep must broadcast rank 0's parameters (seed 0 = the oracle's) before training
    layers = trainer.build_vit_layers(per, embed_dim=dim, heads=4, blocks=2, seed=17 * rank)
    ts = trainer.TrainStep(layers, "vit", (per,) + images.shape[1:], lr=0.02, dp=dp, grad_buckets=3)
    if rank == 0:
        print("gradient all-reduce path:", "NVLink peer-memory kernel" if getattr(dp, "p2p_ready", False) else "NCCL", flush=True)
    worst = 0.0
    for step in range(4):
        ref = M.vit_step(params, images, onehot, 0.02, cfg)
        loss_local = ts.step(images[lo:hi], labels[lo:hi])
        loss = dp.sum_over_ranks(loss_local)          # local losses are already divided by the GLOBAL N
        params = ref["params"]
        mine = M.tree_leaves(trainer.save_walk(layers))
        err = max(float(np.max(np.abs(a.reshape(-1) - b.reshape(-1))) / (np.max(np.abs(b)) + 1e-30))
                  for a, b in zip(mine, M.tree_leaves(params)))
        worst = max(worst, err, abs(loss - ref["loss"]) / abs(ref["loss"]))
        if rank == 0:
            print("step %d loss %.6f (oracle %.6f) max rel param err %.2e graph=%s" % (step, loss, ref["loss"], err, ts.graph is not None), flush=True)
    worst = dp.max_over_ranks(worst)
    if rank == 0:
        fused = any(getattr(l, "_fused", False) for l in layers)
        print("DP_CHECK world=%d dim=%d fused=%s worst=%.2e %s" % (world, dim, fused, worst, "OK" if worst < 1e-4 else "FAIL"), flush=True)
    ts.close()
    return worst


main()
