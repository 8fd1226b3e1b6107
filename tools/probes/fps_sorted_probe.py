"""Bucket-pruned sampling (pn_fps_sorted_f32) against the register-resident kernels: equality and stand-alone time."""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    ts = []
    for i in range(6):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        r = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return statistics.median(ts[1:]), r


for B, N, S, cfgno in ((8, 24000, 1024, 2), (8, 24000, 256, 3), (3, 5000, 300, 4), (2, 33000, 512, 5), (1, 49152, 128, 6), (5, 777, 100, 7),
                       (8, 8000, 1024, 5)):
    x = torch.from_numpy(syn.kitti_batch(B, N, config=cfgno)).to(dev)
    x0 = x.permute(0, 2, 1)[:, :, :3]
    torch.manual_seed(N)
    st = torch.randint(0, N, (B,)).to(dev)
    grid = ops.ball_grid(x0, 0.1)
    t_ref, ref = timed(lambda: ops.fps(x0, S, st))
    t_4, r4 = timed(lambda: ops.fps(x0, S, st, config=(4, 256, 2))) if 8192 <= N <= 24576 else (float("nan"), ref)
    t_new, got = timed(lambda: ops.fps_sorted(x0, grid, S, st))
    t_np, got2 = timed(lambda: ops.fps_sorted(x0, grid, S, st, config=(0, 256, 0)))
    t_8, got3 = timed(lambda: ops.fps_sorted(x0, grid, S, st, config=(0, 512, 0))) if N <= 49152 else (float("nan"), ref)
    t_8np, got4 = timed(lambda: ops.fps_sorted(x0, grid, S, st, config=(0, 512, 1)))
    print(f"B={B} N={N} npoint={S}: auto {t_ref:7.1f} | 4x256 {t_4:7.1f} | pruned32x12 {t_new:7.1f} | pruned8x24 {t_np:7.1f} | pruned16x24 {t_8:7.1f} | "
          f"unpruned16x24 {t_8np:7.1f} us  equal={torch.equal(ref, got) and torch.equal(ref, got2) and torch.equal(ref, got3) and torch.equal(ref, got4)}", flush=True)
