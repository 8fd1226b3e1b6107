import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
x0 = x.permute(0, 2, 1)[:, :, :3]
st = torch.randint(0, 24000, (8,)).to(dev)
grid = ops.ball_grid(x0, 0.1)
ops.fps_sorted(x0, grid, 256, st)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.fps_sorted(x0, grid, 256, st)
ops.fps(x0, 256, st, config=(4, 256, 2))
torch.cuda.synchronize()
torch.cuda.profiler.stop()
