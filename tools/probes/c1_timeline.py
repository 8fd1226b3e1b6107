"""Kernel timeline (CUPTI) of one PointNetSeg forward, B=32 N=1024 (config C4)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402
from pointnet12_b200.model.pointnet import PointNetSeg  # noqa: E402

dev = torch.device("cuda", 0)
net = PointNetSeg(19, input_dims=4, feature_transform=True)
sd = syn.random_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, 1234)
net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
net = net.to(dev).eval()
x = torch.from_numpy(syn.kitti_batch(1, 24000, config=1)).to(dev)
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net(x)
        torch.cuda.synchronize()
evs = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    print(f"{e.time_range.start - t0:8.1f} {e.time_range.end - e.time_range.start:7.1f}  {e.name[:100]}")
print("total", evs[-1].time_range.end - t0)
