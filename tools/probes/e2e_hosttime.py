"""Host-side time per batch of the end-to-end loop: how long submit() and result() take in each output mode."""
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
depth = int(os.environ.get("DEPTH", "6"))
runner = GraphedSemSeg(net, depth=depth)
for mode in (False, True, "labels"):
    for _ in range(2):
        runner.run_pipelined(hs, to_host=mode, consume=lambda k, r: None)
    torch.cuda.synchronize()
    sub, res, pending = [], [], []
    t_all = time.perf_counter()
    for k in range(96):
        t0 = time.perf_counter()
        pending.append(runner.submit(hs[k % 6], to_host=mode))
        t1 = time.perf_counter()
        sub.append(t1 - t0)
        if len(pending) >= depth:
            runner.result(pending.pop(0))
            res.append(time.perf_counter() - t1)
    for t in pending:
        runner.result(t)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t_all) / 96
    print(json.dumps({"to_host": mode, "wall_ms": round(wall * 1e3, 4), "submit_us_median": round(statistics.median(sub) * 1e6, 1),
                      "submit_us_max": round(max(sub) * 1e6, 1), "result_us_median": round(statistics.median(res) * 1e6, 1),
                      "submit_us_mean": round(statistics.mean(sub) * 1e6, 1), "result_us_mean": round(statistics.mean(res) * 1e6, 1)}), flush=True)
