"""fp1's 3-NN search alone (cold L2) on one C2 batch: block search in bucket order vs the all-pairs scan (no order given).
PN12_NN_BLOCK=8|16|32 forces the block size of the search."""
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
from pointnet12_b200.model.pointnet_util import PointNetFeaturePropagation
fp = PointNetFeaturePropagation(128, [128, 128, 128]).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for order in (grid, None):
    ts = []
    for i in range(7):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); r = fp.geometry(x, x1, order=order); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print("block", os.environ.get("PN12_NN_BLOCK", "auto"), "block search" if order is not None else "all-pairs scan", "us", round(statistics.median(ts) * 1e3, 1), "checksum", int(r[0].sum()))
