"""Kernel timeline of CUDA-graph replays of the C2 forward (torch.profiler / CUPTI): start, duration and the idle gap
before every kernel of one replay, in start order.   python tools/graph_timeline.py [replays]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet12_b200 import synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402
from pointnet12_b200.runtime import GraphedSemSeg  # noqa: E402

dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
runner = GraphedSemSeg(net)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    runner(x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(3):
        flush.fill_(i)
        runner(x)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
# last replay = kernels after the last fill
fills = [i for i, e in enumerate(evs) if "FillFunctor" in e.name or "fill" in e.name.lower()]
seq = evs[fills[-1] + 1:]
t0 = seq[0].time_range.start
end_prev = t0
print(f"{'start':>8} {'dur':>7} {'gap':>6}  kernel")
for e in seq:
    s, d = e.time_range.start - t0, e.time_range.end - e.time_range.start
    print(f"{s:8.1f} {d:7.1f} {e.time_range.start - end_prev:6.1f}  {e.name[:70]}")
    end_prev = max(end_prev, e.time_range.end)
print("total", seq[-1].time_range.end - t0)
