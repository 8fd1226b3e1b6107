"""Sweep of the batches-in-flight runner at config C2 (8 clouds x 24000 points): ms per batch, whole region timed with
CUDA events, for combinations of depth / level-1 sampling shape / streamed ball query.  Checks that the pipelined
log-probabilities are torch.equal to the depth-1 ones.

    python tools/pipeline_sweep.py [--steps 40] [--timeline depth]     (knobs: see pointnet12_b200/ops.py, PN12_*)
"""
import argparse
import itertools
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet12_b200 import synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402
from pointnet12_b200.runtime import GraphedSemSeg  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=40)
ap.add_argument("--batches", type=int, default=6)
ap.add_argument("--configs", default="")
ap.add_argument("--timeline", type=int, default=0, help="print a CUPTI kernel timeline of two consecutive batches at this depth")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.cuda.set_device(dev)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
xs = [torch.from_numpy(syn.kitti_batch(8, 24000, config=2, first=8 * i)).to(dev) for i in range(args.batches)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def run(depth, env, steps):
    for k in ("PN12_FPS1", "PN12_STREAM_BALL", "PN12_STREAM_BALL_CTAS", "PN12_STREAM_BALL_SHARE", "PN12_WHATIF", "PN12_NN1_BACKGROUND",
              "PN12_RESERVE_L2", "PN12_FPS1_SORTED", "PN12_BQ_THRESHOLD"):
        os.environ.pop(k, None)
    os.environ.update(env)
    net.module.__dict__.pop("_whatif_cache", None)
    runner = GraphedSemSeg(net, depth=depth)
    torch.manual_seed(7)
    outs = runner.run_pipelined(xs)                       # builds the graphs; results for the equality check
    for _ in range(2):
        runner.run_pipelined(xs, consume=lambda k, r: None)
    torch.cuda.synchronize()
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    runner.run_pipelined((xs[i % len(xs)] for i in range(steps)), consume=lambda k, r: None)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps, outs, runner


base_ms, base_outs, _ = run(1, {}, args.steps)
print(json.dumps({"depth": 1, "env": {}, "ms_per_batch": round(base_ms, 4)}), flush=True)

if args.configs:
    grid = [json.loads(c) for c in args.configs.split(";")]
else:
    grid = []
    for depth, fps1, ball in itertools.product((2, 3), ("", "4,256,2"), ("stream", "after", "share", "few")):
        env = {}
        if fps1:
            env["PN12_FPS1"] = fps1
        if ball == "after":
            env["PN12_STREAM_BALL"] = "0"
        elif ball == "share":
            env["PN12_STREAM_BALL_SHARE"] = "1"
        elif ball == "few":
            env["PN12_STREAM_BALL_CTAS"] = "24"
        grid.append({"depth": depth, "env": env})
for cfg in grid:
    try:
        ms, outs, _ = run(cfg["depth"], cfg["env"], args.steps)
        same = all(torch.equal(o, w) for o, w in zip(outs, base_outs))
        print(json.dumps({**cfg, "ms_per_batch": round(ms, 4), "equal_to_depth1": same}), flush=True)
    except Exception as e:   # noqa: BLE001
        print(json.dumps({**cfg, "error": repr(e)[:300]}), flush=True)

if args.timeline:
    from torch.profiler import ProfilerActivity, profile

    for k in ("PN12_FPS1", "PN12_STREAM_BALL", "PN12_STREAM_BALL_CTAS", "PN12_STREAM_BALL_SHARE"):
        os.environ.pop(k, None)
    if os.environ.get("TIMELINE_ENV"):
        os.environ.update(json.loads(os.environ["TIMELINE_ENV"]))
    runner = GraphedSemSeg(net, depth=args.timeline)
    for _ in range(3):
        runner.run_pipelined(xs, consume=lambda k, r: None)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        runner.run_pipelined((xs[i % len(xs)] for i in range(8)), consume=lambda k, r: None)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    t0 = evs[0].time_range.start
    print(f"{'start':>8} {'dur':>7} {'stream':>6}  kernel")
    for e in evs:
        s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
        if d >= 8.0 or "fps" in e.name:
            print(f"{s:8.1f} {d:7.1f}  {e.name[:90]}")
    print("total", evs[-1].time_range.end - t0)
