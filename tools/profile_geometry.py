"""The level-1 geometry kernels of a C2 forward, each launched alone (for ncu):  ncu --profile-from-start off ... python tools/profile_geometry.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
x0 = x.permute(0, 2, 1)[:, :, :3]
torch.manual_seed(0)
st = torch.randint(0, 24000, (8,)).to(dev)
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    fps1 = ops.fps(x0, 1024, st, config=(4, 256, 2) if i == 2 else None)
    x1 = ops.index_points(x0, fps1)
    grid = ops.ball_grid(x0, 0.1)
    ball = ops.ball_query(0.1, 32, x0, x1, grid=grid)
    nn = ops.three_nn(x0, x1, order=grid)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
