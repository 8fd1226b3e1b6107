"""Level-1 ball query alone (cold L2) on one C2 batch, for a few candidate thresholds of the cell path."""
import os, sys, statistics
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev).permute(0, 2, 1)[:, :, :3]
torch.manual_seed(0)
st = torch.randint(0, 24000, (8,)).to(dev)
x1 = ops.index_points(x, ops.fps(x, 1024, st))
grid = ops.ball_grid(x, 0.1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for thr in (None, 1280, 1920, 2560, 3840, 5120, 8192, 2 ** 31 - 1):
    ts = []
    for i in range(7):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = ops.ball_query(0.1, 32, x, x1, grid=grid, threshold=thr); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("ball query level 1 threshold", thr, "us", round(statistics.median(ts) * 1e3, 1), "checksum", int(r.sum()))
