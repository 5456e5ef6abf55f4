#!/usr/bin/env python
"""bench.py — the headline measurement (BASELINE.json): train images/s of the ViT-MNIST step
(forward + backward + update), B200 vs host NumPy.

  python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU under torchrun)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port), rank 0 only

One "step" = forward of every layer + cross_entropy_loss + backward + update + setzero for one
batch (transformer_of_image.py:141-159).  Workloads: vit_readme (default, the config the metric is
quoted on: B=100/GPU, D=96, 6x4 heads), vit_scaled, gpt_poetry, gpt_scaled, gpt_char.
Data parallel runs are weak-scaled: every rank processes a full per-GPU batch, gradients are
all-reduced over NCCL, value = global units / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (kind, dict)  — shapes from BASELINE.json configs / SURVEY.md §8
    "vit_readme": ("vit", dict(batch=100, embed_dim=96, heads=4, blocks=6, lr=1e-3, adam=False)),
    "vit_scaled": ("vit", dict(batch=1024, embed_dim=384, heads=6, blocks=12, lr=1e-3, adam=False)),
    "gpt_poetry": ("gpt", dict(batch=64, context=256, embed_dim=384, heads=6, blocks=5, vocab=2350, lr=3e-4, adam=True)),
    "gpt_scaled": ("gpt", dict(batch=64, context=256, embed_dim=768, heads=12, blocks=12, vocab=2350, lr=3e-4, adam=True)),
    "gpt_char": ("gpt", dict(batch=2, context=5, embed_dim=192, heads=3, blocks=1, vocab=26, lr=3e-3, adam=True)),
}
METRIC = "train images/s (ViT-MNIST fwd+bwd+update)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf_burst=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    source="measured")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback")


def step_flops(kind, c):
    """Algorithmic FLOPs per step (GEMM + attention, bwd = 2x fwd), SURVEY.md §8(d)."""
    D, H, L, B = c["embed_dim"], c["heads"], c["blocks"], c["batch"]
    if kind == "vit":
        N, P, C = 49, 16, 10
        M = B * N
        lin = 2 * M * (P * D + L * (D * 3 * D + D * 2 * D + 2 * D * D)) + 2 * B * (N * D * D + D * C)
        att = L * B * H * 4 * N * N * (D // H)
    else:
        T, V = c["context"], c["vocab"]
        M = B * T
        lin = 2 * M * (L * (D * 3 * D + 2 * D * 4 * D + D * D) + D * D + D * V)
        att = L * B * H * 4 * T * T * (D // H)
    return 3 * (lin + att)


class ClockSampler(threading.Thread):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm = sorted(int(r[1]) for r in self.rows if len(r) > 2 and r[1].isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        smmax = max([int(r[2]) for r in self.rows if len(r) > 2 and r[2].isdigit()] or [0])
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=smmax or None, reasons=sorted(reasons),
                    samples=len(sm))


def op_cost(name, a, kind, c):
    """(flops, bytes) of one profiled C-ABI op from its integer arguments."""
    if name in ("nt_linear_fwd", "nt_linear_bwd_input", "nt_linear_bwd_params"):
        M, K, N = a[0], a[1], a[2]
        return 2.0 * M * K * N, 4.0 * (M * K + K * N + M * N)
    if name in ("nt_linear_fwd_bf16x3", "nt_linear_bwd_input_bf16x3"):
        M, K, N = a[0], a[1], a[2]
        return 2.0 * M * K * N, 4.0 * (M * K + K * N + M * N)
    if name == "nt_linear_bwd_params_bf16x3":
        M, K, N = a[-3:]
        return 2.0 * M * K * N, 4.0 * (M * K + K * N + M * N)
    if name == "nt_split_bf16":
        rows, cols = a[0], a[1]
        return 0.0, 8.0 * rows * cols                       # read fp32, write bf16 hi + lo
    if name == "nt_split_bf16_both":
        rows, cols = a[0], a[1]
        return 0.0, 12.0 * rows * cols                      # read fp32 once, write both layouts
    if name in ("nt_vit_blocks_fwd", "nt_vit_blocks_bwd"):
        # fused ViT block stack: L blocks x B images; per image and block the three Linear layers (7 D^2 weights)
        # and the attention core (QK^T and PV over N tokens); backward = 2x forward
        L, B, N, D, H = a[-5:]
        fwd = L * B * (2.0 * N * 7 * D * D + 4.0 * N * N * D)
        act = 4.0 * L * B * N * (3 * D + H * N + D + 2 * D + D)      # qkv, P, h, dgelu, out
        return (fwd, act) if name.endswith("fwd") else (2.0 * fwd, act)
    if name == "nt_attention_fwd":
        B, N, H, dh = a[-4:]
        return 4.0 * B * H * N * N * dh, 4.0 * (4 * B * N * H * dh + B * H * N * N)
    if name == "nt_attention_bwd":
        B, N, H, dh = a[-5:-1]                  # trailing int is the mask kind
        return 8.0 * B * H * N * N * dh, 4.0 * (7 * B * N * H * dh + B * H * N * N)
    if name == "nt_layernorm_fwd":
        M, D = a[0], a[1]
        return 0.0, 8.0 * M * D
    if name == "nt_layernorm_bwd":
        M, D = a[0], a[1]
        return 0.0, 16.0 * M * D
    if name == "nt_cross_entropy":
        r, cc = a[-2], a[-1]
        return 0.0, 12.0 * r * cc
    if name == "nt_sgd":
        return 0.0, 16.0 * a[0]
    if name == "nt_adam_star":
        return 0.0, 32.0 * a[0]
    return 0.0, 0.0


def run_ours(args):
    import numpy_transformer_b200 as nt
    from numpy_transformer_b200 import trainer
    from numpy_transformer_b200.parallel import DataParallel

    kind, c = WORKLOADS[args.workload]
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with: python -m torch.distributed.run --nproc-per-node %d bench.py --gpus %d ..." % (args.gpus, args.gpus))
    dp = DataParallel(rank, world) if world > 1 else None
    nt.init(local)
    if args.gemm_mode is not None:
        nt.set_gemm_mode(args.gemm_mode)
    if dp:
        dp.init_nccl()
    B = c["batch"]
    chains = args.chains if args.chains is not None else 1
    if kind == "vit":
        reps = [trainer.build_vit_layers(B // chains, embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"], seed=0)
                for _ in range(chains)]
        layers = reps[0]
        rs = np.random.RandomState(1 + rank)
        inputs = rs.rand(B, 1, 28, 28)
        labels = rs.randint(0, 10, B)
        units, unit = B, "images/s"
    else:
        reps = [trainer.build_gpt_layers(B // chains, c["context"], c["embed_dim"], c["heads"], c["blocks"], c["vocab"], adam=True, seed=0)
                for _ in range(chains)]
        layers = reps[0]
        rs = np.random.RandomState(1 + rank)
        inputs = rs.randint(0, c["vocab"], (B, c["context"]))
        labels = rs.randint(0, c["vocab"], (B, c["context"]))
        units, unit = B, "sequences/s"
    ts = trainer.TrainStep(layers, kind, inputs.shape, lr=c["lr"], adam=c["adam"], dp=dp, grad_buckets=args.buckets,
                           replicas=reps[1:])

    W = max(3, args.warmup)
    K = args.steps
    for _ in range(W):
        ts.step(inputs, labels)
    from numpy_transformer_b200._lib import lib
    import ctypes

    # ---------------- device-timed region: inputs resident in HBM, L2 flushed between timed steps
    ev = []
    for _ in range(2 * K):
        e = ctypes.c_void_p()
        lib.nt_event_create(ctypes.byref(e))
        ev.append(e)
    ts.load(inputs, labels)
    lib.nt_sync()
    if dp:
        dp.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = nt.launch_count()
    t_wall0 = time.perf_counter()
    for i in range(K):
        lib.nt_flush_l2()
        lib.nt_event_record(ev[2 * i])
        ts.run_device()
        lib.nt_event_record(ev[2 * i + 1])
    lib.nt_sync()
    t_wall = time.perf_counter() - t_wall0
    launches = nt.launch_count() - launches0
    if dp:
        dp.barrier()
    ms = ctypes.c_float(0)
    per_step = []
    for i in range(K):
        lib.nt_event_elapsed_ms(ev[2 * i], ev[2 * i + 1], ctypes.byref(ms))
        per_step.append(ms.value)
    dev_s = sum(per_step) / 1e3
    dev_s_max = dp.max_over_ranks(dev_s) if dp else dev_s
    value = units * world * K / dev_s_max

    # ---------------- end-to-end through the public call: pinned host batch in, host loss out
    lib.nt_sync()
    if dp:
        dp.barrier()
    # the synthetic batch sits in the step's pinned staging buffers (a loader with pinned output buffers); every
    # timed step copies it host -> device from there, runs the step and reads the loss back
    ts.h_in.array[...] = np.asarray(inputs).reshape(ts.input_shape)
    ts.h_lab.array[...] = np.asarray(labels).reshape(-1)
    t0 = time.perf_counter()
    for _ in range(K):
        loss = ts.step()
    e2e_s = time.perf_counter() - t0
    e2e_s_max = dp.max_over_ranks(e2e_s) if dp else e2e_s
    clocks = sampler.stop()
    e2e = dict(value=units * world * K / e2e_s_max, unit=unit, h2d_bytes_per_step=ts.h2d_bytes, d2h_bytes_per_step=4 * chains,
               timing="host wall clock around K public TrainStep.step() calls (batch in pinned host memory -> H2D -> step -> D2H loss), max over ranks")

    # ---------------- roofline of the dominant kernel: the same step captured with a CUDA-event pair around every op
    # (event-record graph nodes), replayed; durations are device times with no host gaps between ops
    pk = peaks()
    prof_steps = 5
    pg, precs = ts.capture_profiled()
    recs = []
    for _ in range(prof_steps):
        lib.nt_flush_l2()
        if dp is not None:
            dp.barrier()                # ranks enter the replay together: the all-reduce wait is not charged with their skew
        lib.nt_graph_launch(pg)
        lib.nt_sync()
        recs += lib.profile_read(precs)
    ts.steps_done += prof_steps
    agg = {}
    for name, a, ms_ in recs:
        key = (name, a)
        r = agg.setdefault(key, [0.0, 0])
        r[0] += ms_
        r[1] += 1
    total_ms = sum(v[0] for v in agg.values())
    classes = {}
    for (name, a), (tms, n) in agg.items():
        fl, by = op_cost(name, a, kind, c)
        cl = classes.setdefault(name, dict(ms=0.0, n=0, flops=0.0, bytes=0.0))
        cl["ms"] += tms
        cl["n"] += n
        cl["flops"] += fl * n
        cl["bytes"] += by * n
    # the dominant KERNEL: waits on the collective (nt_dp_*) stay in the op table but are not a kernel of this repo
    top = max(((k, v) for k, v in classes.items() if not k.startswith("nt_dp_")), key=lambda kv: kv[1]["ms"])
    tname, tc = top
    if tc["flops"] > 0:
        ach = tc["flops"] / (tc["ms"] / 1e3) / 1e12
        roof = dict(bound="tensor", kernel=tname, achieved=ach, peak=pk["tf_burst"], unit="TFLOP/s", frac=ach / pk["tf_burst"],
                    traffic=None, share_of_step=tc["ms"] / total_ms, launches_per_step=tc["n"] / prof_steps,
                    peak_source=pk["source"] + " bf16 burst")
    else:
        ach = tc["bytes"] / (tc["ms"] / 1e3) / 1e9
        roof = dict(bound="hbm", kernel=tname, achieved=ach, peak=pk["hbm"], unit="GB/s", frac=ach / pk["hbm"],
                    traffic=None, share_of_step=tc["ms"] / total_ms, launches_per_step=tc["n"] / prof_steps,
                    peak_source=pk["source"] + " copy")
    # traffic: DRAM bytes of the dominant op from the committed ncu capture (null when none was taken for this workload)
    try:
        roof["traffic"] = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload, {}).get(tname)
    except Exception:
        pass
    if tname.startswith("nt_vit_blocks"):
        # the fused blocks execute every product as three BF16 MMA passes on 64 padded rows per 49-token image
        roof["executed_mma_tflops"] = ach * 3.0 * 64.0 / 49.0
        roof["note"] = ("algorithmic FLOPs / measured time vs the dense BF16 tcgen05 peak; the kernel issues 3 BF16 mma.sync passes "
                        "per product (fp32-class accuracy) on 64/49 padded rows with one image per CTA (B of 148 SMs busy); the "
                        "legacy mma.sync path peaks at 556 TFLOP/s chip-wide (tools/mma_red_bench.cu)")
    fl_step = step_flops(kind, c)
    step_tf = fl_step * world / (dev_s_max / K) / 1e12
    op_table = sorted(((n, round(v["ms"] / prof_steps, 4), v["n"] // prof_steps,
                        round(v["flops"] / max(v["ms"], 1e-9) / 1e9, 3), round(v["bytes"] / max(v["ms"], 1e-9) / 1e6, 2))
                       for n, v in classes.items()), key=lambda r: -r[1])

    out = None
    if rank == 0:
        out = {
            "metric": METRIC if kind == "vit" else "train sequences/s (GPT fwd+bwd+update)",
            "value": value, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dev_s_max / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "per_gpu_batch": B, "global_batch": B * world, **{k: v for k, v in c.items() if k not in ("batch",)},
                       "parallelism": "dp%d" % world, "micro_batch_chains_per_gpu": chains, "gemm_mode": nt.get_gemm_mode(),
                       "l2": "256 MB memset between timed steps (working set is L2-resident)",
                       "graph_kernels_per_step": ts.graph_kernels, "params": ts.arena.n_params},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof,
            "step_roofline": {"flops_per_step": fl_step, "achieved_tflops": step_tf, "frac_of_bf16_sustained": step_tf / (pk["tf_sustained"] * world)},
            "op_table_ms_per_step(name,ms,launches,GFLOP/s,GB/s)": op_table[:12],
            "loss": loss,
        }
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_budget)
    lib.nt_graph_destroy(pg)
    ts.close()
    if dp:
        dp.shutdown()
    return out


def cpu_baseline(workload, budget_s=20.0, refstruct_too=True):
    """Oracle port of the reference step timed on this host (all BLAS threads NumPy uses)."""
    from oracle import models as M
    kind, c = WORKLOADS[workload]
    try:
        from threadpoolctl import threadpool_info
        cores = max([i.get("num_threads", 1) for i in threadpool_info()] or [os.cpu_count()])
    except Exception:
        cores = os.cpu_count()
    if kind != "vit":
        return cpu_baseline_gpt(workload, budget_s, cores)
    B = min(c["batch"], 100)
    cfg = M.ViTConfig(batch=B, embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"])
    params = M.vit_init(cfg, 0)
    images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
    M.vit_step(params, images, onehot, c["lr"], cfg)
    n, t0 = 0, time.perf_counter()
    while True:
        out = M.vit_step(params, images, onehot, c["lr"], cfg)
        params = out["params"]
        n += 1
        if time.perf_counter() - t0 > budget_s * 0.5 or n >= 20:
            break
    fast = B * n / (time.perf_counter() - t0)
    res = dict(value=fast, unit="images/s", cores=cores, kind="port",
               sample="%d steps of batch %d of %s (closed-form softmax backward, vectorised over batch x heads)" % (n, B, workload))
    if refstruct_too:
        Bs = min(B, 20)
        cfg2 = M.ViTConfig(batch=Bs, embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"])
        im2, _, oh2 = M.synthetic_vit_batch(cfg2, 1)
        p2 = M.vit_init(cfg2, 0)
        t0 = time.perf_counter()
        M.vit_step(p2, im2, oh2, c["lr"], cfg2, refstruct=True)
        res["reference_structured"] = dict(value=Bs / (time.perf_counter() - t0), unit="images/s",
                                           sample="1 step of batch %d with the reference's per-sample/per-head loops and per-row softmax Jacobian (no ThreadPool churn)" % Bs)
    return res


def cpu_baseline_gpt(workload, budget_s, cores):
    from oracle import models as M
    kind, c = WORKLOADS[workload]
    B = min(c["batch"], 2)
    cfg = M.GPTConfig(batch=B, context=c["context"], embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"], vocab=c["vocab"])
    params = M.gpt_init(cfg, 0)
    opt = M.adam_state(params)
    tokens, labels = M.synthetic_gpt_batch(cfg, 1)
    t0 = time.perf_counter()
    M.gpt_step(params, opt, tokens, labels, c["lr"], cfg)
    return dict(value=B / (time.perf_counter() - t0), unit="sequences/s", cores=cores, kind="port",
                sample="1 step of batch %d of %s" % (B, workload))


def run_reference(args):
    """The reference's own CPU path (oracle port: the Python reference cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import models as M
    kind, c = WORKLOADS[args.workload]
    world = args.gpus
    W, K = max(1, min(args.warmup, 2)), max(1, min(args.steps, 10))
    if kind == "vit":
        B = min(c["batch"], 100)
        cfg = M.ViTConfig(batch=B, embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"])
        params = M.vit_init(cfg, 0)
        images, labels, onehot = M.synthetic_vit_batch(cfg, 1)
        for _ in range(W):
            params = M.vit_step(params, images, onehot, c["lr"], cfg)["params"]
        t0 = time.perf_counter()
        for _ in range(K):
            params = M.vit_step(params, images, onehot, c["lr"], cfg)["params"]
        dt = time.perf_counter() - t0
        unit, metric = "images/s", METRIC
    else:
        B = min(c["batch"], 2)
        cfg = M.GPTConfig(batch=B, context=c["context"], embed_dim=c["embed_dim"], heads=c["heads"], blocks=c["blocks"], vocab=c["vocab"])
        params, tokens_labels = M.gpt_init(cfg, 0), M.synthetic_gpt_batch(cfg, 1)
        opt = M.adam_state(params)
        K, W = min(K, 2), 0
        t0 = time.perf_counter()
        for _ in range(K):
            o = M.gpt_step(params, opt, tokens_labels[0], tokens_labels[1], c["lr"], cfg)
            params, opt = o["params"], o["opt"]
        dt = time.perf_counter() - t0
        unit, metric = "sequences/s", "train sequences/s (GPT fwd+bwd+update)"
    try:
        from threadpoolctl import threadpool_info
        cores = max([i.get("num_threads", 1) for i in threadpool_info()] or [os.cpu_count()])
    except Exception:
        cores = os.cpu_count()
    v = B * K / dt
    sample = "%d steps of batch %d of %s on the host CPU (oracle port of the reference, closed-form softmax backward)" % (K, B, args.workload)
    return {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": dt / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": args.workload, "per_gpu_batch": B, **{k: v_ for k, v_ in c.items() if k != "batch"}},
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="vit_readme", choices=sorted(WORKLOADS))
    ap.add_argument("--gemm-mode", type=int, default=None)
    ap.add_argument("--buckets", type=int, default=1)
    ap.add_argument("--chains", type=int, default=None, help="parallel micro-batch branches per GPU (default 1; measured slower than 1 on B200 for vit_readme, see DESIGN.md)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    args = ap.parse_args()
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out is not None:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
