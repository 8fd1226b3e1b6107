"""Grid-size sweep of the stand-alone level-1 ball query and the fp1 3-NN block search (env caps PN12_BQ_GX / PN12_NN_GX)."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
x0 = x.permute(0, 2, 1)[:, :, :3]
torch.manual_seed(0)
fps1 = ops.fps(x0, 1024, torch.randint(0, 24000, (8,)).to(dev))
x1 = ops.index_points(x0, fps1)
grid = ops.ball_grid(x0, 0.1)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for i in range(7):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[2:]), r


ref_ball = ops.ball_query(0.1, 32, x0, x1, grid=grid)
ref_nn = ops.three_nn(x0, x1, order=grid)
for cap in (0, 74, 55, 37, 28, 19):
    os.environ["PN12_BQ_GX"] = str(cap)
    t, r = timed(lambda: ops.ball_query(0.1, 32, x0, x1, grid=grid))
    print(f"ball query grid.x cap {cap:4d} (x8 clouds): {t:7.1f} us  equal={torch.equal(r, ref_ball)}")
os.environ["PN12_BQ_GX"] = "0"
for cap in (0, 185, 148, 111, 94, 74, 63, 56, 47, 37):
    os.environ["PN12_NN_GX"] = str(cap)
    t, r = timed(lambda: ops.three_nn(x0, x1, order=grid))
    print(f"3-NN grid.x cap {cap:4d} (x8 clouds): {t:7.1f} us  equal={torch.equal(r[0], ref_nn[0]) and torch.equal(r[1], ref_nn[1])}")
