"""Does pacing the submits (minimum spacing f x the running time per batch) break the convoy of the end-to-end loop?"""
import json
import os
import statistics
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402
from pointnet12_b200.runtime import GraphedSemSeg  # noqa: E402

dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
hs = [torch.from_numpy(syn.kitti_batch(8, 24000, config=2, first=8 * i)).pin_memory() for i in range(6)]
ds = [h.to(dev) for h in hs]
STEPS = 192


def loop(runner, depth, batches, mode, f, fixed=None):
    pending, res = [], []
    t_est, t_prev_done, t_last = None, None, 0.0
    t_all = time.perf_counter()
    for k in range(STEPS):
        spacing = fixed if fixed is not None else (f * t_est if (f and t_est) else 0.0)
        while time.perf_counter() < t_last + spacing:
            pass
        t_last = time.perf_counter()
        pending.append(runner.submit(batches[k % 6], to_host=mode))
        if len(pending) >= depth:
            t1 = time.perf_counter()
            tk = pending.pop(0)
            if mode is False:
                tk.done.synchronize()
            runner.result(tk)
            now = time.perf_counter()
            res.append(now - t1)
            if t_prev_done is not None:
                dt = now - t_prev_done
                t_est = dt if t_est is None else 0.9 * t_est + 0.1 * dt
            t_prev_done = now
    for t in pending:
        runner.result(t)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_all) / STEPS
    return wall, res


for depth in (4, 6):
    runner = GraphedSemSeg(net, depth=depth)
    for mode, batches in ((False, ds), (True, hs), ("labels", hs)):
        runner.run_pipelined(batches, to_host=mode, consume=lambda k, r: None)
        for f, fixed in ((0.0, None), (0.8, None), (0.9, None), (0.97, None), (None, 0.00036), (None, 0.00039), (None, 0.00042)):
            loop(runner, depth, batches, mode, f, fixed)
            wall, res = loop(runner, depth, batches, mode, f, fixed)
            print(json.dumps({"depth": depth, "to_host": mode, "f": f, "fixed_us": None if fixed is None else fixed * 1e6,
                              "wall_ms": round(wall * 1e3, 4), "blocked_frac": round(sum(r > 50e-6 for r in res) / len(res), 3),
                              "result_us_max": round(max(res) * 1e6)}), flush=True)
