"""Config C5 (BASELINE.json configs[4]): PointNet2SemSeg training step, batch 8 x 8000 points per GPU, data-parallel.

    python tools/bench_train.py [--steps 20 --warmup 5 --batch 8 --points 8000]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_train.py

A step = forward in train() mode (batch-statistics BatchNorm, dropout) + the reference's loss + backward + gradient
all-reduce (NCCL, one flat 3.9 MB buffer; skipped at N=1) + Adam, all through the package's public training API
(pointnet12_b200.train).  Inputs resident; CUDA events around every step; MAX over ranks.  Prints one JSON line
(rank 0) with points/s, ms/step, e2e, clocks and the live roofline of the dominant kernel.  The CPU leg (`cpu_baseline`: the
float64 oracle of the iteration) is supplied by bench.py -- `python bench.py --workload train` -- because only bench.py's CPU legs
may execute oracle/; run directly, this tool reports "cpu_baseline": null.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main(argv=None, cpu_baseline=None, init_pg=True, emit=True):
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--points", type=int, default=8000)
    ap.add_argument("--eager", action="store_true", help="launch kernel by kernel through autograd instead of one CUDA-graph replay")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--model", default="pointnet2", choices=["pointnet2", "pointnet"],
                    help="pointnet2: PointNet2SemSeg (config C5, default); pointnet: PointNetSeg(19, 4, feature_transform=True), the "
                         "default model of the reference's driver, eager launches, loss + 0.001 * feature_transform_reguliarzer")
    args = ap.parse_args(argv)
    import torch.distributed as dist

    from bench import ClockSampler

    from pointnet12_b200 import _native as nv, ops, synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.train import FlatAdam, GraphedTrainStep, cross_entropy

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1 and init_pg:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(1234)                                   # same initial weights on every rank (DataParallel replicas)
    if args.model == "pointnet":
        from pointnet12_b200.model.pointnet import PointNetSeg, feature_transform_reguliarzer

        net = PointNetSeg(19, input_dims=4, feature_transform=True).to(dev).train()
    else:
        net = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    opt = FlatAdam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    B, N = args.batch, args.points
    pts = torch.from_numpy(syn.kitti_batch(B, N, config=5, first=rank * B)).to(dev)
    target = torch.from_numpy(np.random.default_rng(5000 + rank).integers(0, 19, size=(B, N))).to(dev)
    torch.manual_seed(rank)

    def step(marks=None):
        def mark(name):
            if marks is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                marks.append((name, e))
        mark("start")
        if args.model == "pointnet":
            logp, trans_feat = net(pts)
            mark("forward")
            loss = cross_entropy(logp, target) + feature_transform_reguliarzer(trans_feat) * 0.001
        else:
            logp = net(pts)
            mark("forward")
            loss = cross_entropy(logp, target)
        mark("loss")
        opt.zero_grad()
        loss.backward()
        mark("backward")
        scale = opt.all_reduce()
        mark("allreduce")
        opt.step(scale)
        mark("adam")
        return loss

    eager_step = step
    if args.eager:
        graphed = None
    elif args.model == "pointnet":
        from pointnet12_b200.train import GraphedStep

        def pn_loss(n, x, t):
            out, trans_feat = n(x)
            return cross_entropy(out, t) + feature_transform_reguliarzer(trans_feat) * 0.001

        gstep = GraphedStep(net, opt, pn_loss)
        graphed = lambda x, t, next_points=None: gstep(x, t)   # noqa: E731  (nothing to prefetch: no sampling stage)
    else:
        graphed = GraphedTrainStep(net, opt)
    if graphed is not None:
        def step(marks=None):                                  # noqa: F811
            # the upcoming batch is known (a loader runs ahead): its geometry is prefetched during this iteration
            return graphed(pts, target, next_points=pts) if marks is None else eager_step(marks)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    n0 = nv.launch_count
    evs = []
    with ClockSampler(local) as clocks:
        for _ in range(args.steps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            loss = step()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
    launches = nv.launch_count - n0
    # end to end: the batch starts in pinned host memory every step and the loss is read back
    host_pts, host_tgt = pts.cpu().pin_memory(), target.cpu().pin_memory()
    host_loss = torch.empty((), dtype=torch.float32).pin_memory()
    e2e = []
    dbuf = [pts, pts.clone()]
    for i in range(args.steps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        cur, nxt = dbuf[i & 1], dbuf[1 - (i & 1)]
        nxt.copy_(host_pts, non_blocking=True)              # the UPCOMING batch arrives while this one is processed
        target.copy_(host_tgt, non_blocking=True)
        if graphed is not None:
            host_loss.copy_(graphed(cur, target, next_points=nxt), non_blocking=True)
        else:
            pts.copy_(host_pts, non_blocking=True)
            host_loss.copy_(step(), non_blocking=True)
        b.record()
        e2e.append((a, b))
    torch.cuda.synchronize()
    e2e_total = torch.tensor([sum(a.elapsed_time(b) for a, b in e2e)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_total, op=dist.ReduceOp.MAX)
    times = np.array([a.elapsed_time(b) for a, b in evs])
    total = torch.tensor([times.sum()], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total, op=dist.ReduceOp.MAX)
    # roofline of the dominant kernel (the fused layer GEMM, HBM-bound): CUDA events around every call of one eager
    # iteration; algorithmic bytes = rows * (cin + cout) * 4 per call (operand read once, output written once)
    roofline = None
    if ops.mlp_mode() == "bf16x3":
        nv.time_entry_points(["pn_train_gemm_bf16x3"])
        for _ in range(3):
            flush.fill_(1)
            eager_step()
        torch.cuda.synchronize()
        recs = nv.time_entry_points(None)["pn_train_gemm_bf16x3"]
        if recs:
            n_calls = len(recs) // 3
            recs = recs[-n_calls:]                            # the last of the three iterations
            t_ms = sum(a.elapsed_time(b) for a, b, _ in recs)
            nbytes = sum(r * (ci + co) * 4 for _, _, (r, ci, co) in recs)
            try:
                peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                peak, src = float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
            except Exception:
                peak, src = 6650.0, "fallback (B200_PROFILING.md)"
            ach = nbytes / (t_ms * 1e-3) / 1e9
            roofline = {"kernel": "gemm::train_gemm_kernel (pn_train_gemm_bf16x3: forward layer GEMMs with fused BatchNorm, input-gradient GEMMs)",
                        "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                        "peak_source": src, "algorithmic_bytes_per_step": int(nbytes), "launches_per_step": n_calls,
                        "kernel_ms_per_step": t_ms, "share_of_step": t_ms / (float(total.item()) / args.steps),
                        "timing": "CUDA events around each call in an eager iteration after the timed region (includes ~2 us of launch gap per call)"}
    # the gradient exchange alone: one NCCL all-reduce of the flat 3.9 MB gradient (device-timed, mean of 20, max over ranks)
    allreduce_us = None
    if world > 1:
        for _ in range(3):
            opt.all_reduce()
        torch.cuda.synchronize()
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(20):
            opt.all_reduce()
        b.record()
        torch.cuda.synchronize()
        t_ar = torch.tensor([a.elapsed_time(b) / 20 * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(t_ar, op=dist.ReduceOp.MAX)
        allreduce_us = float(t_ar.item())
    phases = None
    if args.eager:                       # per-phase split of one more (instrumented) eager iteration
        marks = []
        step(marks)
        torch.cuda.synchronize()
        phases = {n1: round(e0.elapsed_time(e1), 4) for (n0_, e0), (n1, e1) in zip(marks[:-1], marks[1:])}
    record = None
    if rank == 0:
        ms = float(total.item()) / args.steps
        record = {
            "metric": "pointnet2_semseg_train_points_per_sec", "value": world * B * N / (ms * 1e-3), "unit": "points/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "dtype": "f32 (GEMMs as 3-pass split bf16 on tensor cores)" if ops.mlp_mode() == "bf16x3" else "f32 (CUDA-core GEMMs)", "data": "synthetic",
            "config": {"workload": f"{'PointNetSeg(19, 4, feature_transform=True)' if args.model == 'pointnet' else 'C5: PointNet2SemSeg(19, feature_dims=1)'} training step (forward, CrossEntropyLoss, backward, "
                                   f"gradient all-reduce, Adam), {B} clouds x {N} points per GPU, seeded random init",
                       "l2": "256 MiB written between timed steps", "launch": "eager" if args.eager else "forward + loss + backward (+ the next batch's sampling / grouping on a side stream) as one CUDA-graph replay, then all-reduce and Adam"},
            "step_ms": {"min": float(times.min()), "median": float(np.median(times)), "max": float(times.max())},
            "vs_baseline": None, "clocks": clocks.summary(), "roofline": roofline,
            "e2e": {"value": world * B * N / (float(e2e_total.item()) / args.steps * 1e-3), "unit": "points/s",
                    "ms_per_step": float(e2e_total.item()) / args.steps,
                    "h2d_bytes_per_step": int(host_pts.numel() * 4 + host_tgt.numel() * 8), "d2h_bytes_per_step": 4},
            "cpu_baseline": None if (args.no_cpu_baseline or world > 1 or cpu_baseline is None) else cpu_baseline(),
            "allreduce_us": allreduce_us, "allreduce_bytes": int(opt.grad.numel() * 4),
            "phases_ms_eager": phases, "gpu_launches": int(launches), "final_loss": float(loss.item())}
        if emit:
            print(json.dumps(record))
    if world > 1 and init_pg:
        dist.destroy_process_group()
    return record


def dp_check(world: int, rank: int, dev, points: int = 8000, clouds: int = 8):
    """Data-parallel equivalence on hardware (the reference's DataParallel, pcdseg.py:141):
    a fixed global batch of 8 * world clouds is processed (a) sharded, 8 clouds per rank, and (b) by rank 0 alone, shard after
    shard.  Eval forward: the per-cloud log-probabilities must be torch.equal (clouds are independent).  Training iteration
    (seeded weights, BatchNorm statistics per replica as DataParallel computes them, same dropout streams): the all-reduced
    mean gradient against the mean of rank 0's per-shard gradients, relative L2 (atomics reorder sums: <= 1e-5)."""
    import torch.distributed as dist

    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.train import cross_entropy, semseg_forward_train

    def shard(s):
        x = torch.from_numpy(syn.kitti_batch(clouds, points, config=2, first=5000 + clouds * s)).to(dev)
        g = torch.Generator().manual_seed(777 + s)
        st = [torch.randint(0, n, (clouds,), generator=g, dtype=torch.long).to(dev) for n in (points, 1024, 256, 64)]
        tgt = torch.randint(0, 19, (clouds, points), generator=g, dtype=torch.long).to(dev)
        so = torch.tensor([4242 + s, 1], dtype=torch.int64, device=dev)          # dropout stream {seed, offset} of this shard
        return x, st, tgt, so

    torch.manual_seed(1234)                                   # identical replicas
    net = PointNet2SemSeg(19, feature_dims=1).to(dev)
    params = [p for p in net.parameters() if p.requires_grad]

    def eval_logp(s):
        x, st, _, _ = shard(s)
        net.eval()
        with torch.no_grad():
            return net(x, fps_starts=st)

    def train_grad(s):
        x, st, tgt, so = shard(s)
        net.train()
        for p in params:
            p.grad = None
        loss = cross_entropy(semseg_forward_train(net, x, st, seed_offset=so), tgt)
        loss.backward()
        return torch.cat([p.grad.reshape(-1) for p in params]).clone()

    state = {k: v.clone() for k, v in net.state_dict().items()}   # train() forwards move the BatchNorm running statistics

    def reset():
        net.load_state_dict(state)

    mine = eval_logp(rank)
    g_mine = train_grad(rank)
    reset()
    if world > 1:
        gathered = [torch.empty_like(mine) for _ in range(world)] if rank == 0 else None
        dist.gather(mine, gathered, dst=0)
        g_mean = g_mine.clone()
        dist.all_reduce(g_mean, op=dist.ReduceOp.SUM)
        g_mean /= world
    else:
        gathered, g_mean = [mine], g_mine
    if rank != 0:
        return None
    equal = True
    g_alone = torch.zeros_like(g_mean)
    for s in range(world):
        equal = equal and bool(torch.equal(eval_logp(s), gathered[s]))
        g_alone += train_grad(s)
        reset()
    g_alone /= world
    rel = float((g_mean - g_alone).norm() / g_alone.norm())
    # noise floor: the SAME shard twice on the same GPU (the backward kernels accumulate with fp32 atomics, whose order varies)
    g_a = train_grad(0)
    reset()
    g_b = train_grad(0)
    reset()
    floor = float((g_a - g_b).norm() / g_a.norm())
    return {"global_batch_clouds": clouds * world, "points_per_cloud": points, "world": world,
            "logp_equal_across_world_sizes": equal, "mean_gradient_rel_l2": rel, "rerun_noise_floor_rel_l2": floor,
            "gradient_ok": rel <= max(1e-5, 3.0 * floor),
            "what": "the same global batch sharded over the ranks vs processed by rank 0 alone, shard by shard (BatchNorm statistics "
                    "per replica, as DataParallel computes them); eval log-probs compared with torch.equal, the all-reduced mean "
                    "gradient of one training iteration by relative L2 (bar: 1e-5, or 3x the run-to-run floor of one GPU repeating "
                    "one shard -- fp32 atomics in the backward kernels reorder sums)"}


if __name__ == "__main__":
    main()
