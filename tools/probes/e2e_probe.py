"""End-to-end (pinned host in, pinned host out) ms per batch of the batches-in-flight runner for a few knobs."""
import json
import os
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
DEPTHS = [int(d) for d in os.environ.get("E2E_DEPTHS", "4,6").split(",")]
SLICES = os.environ.get("E2E_SLICES", "1,2,8").split(",")
for depth in DEPTHS:
    for slices in SLICES:
        os.environ["PN12_PIPE_HOST_SLICES"] = slices
        runner = GraphedSemSeg(net, depth=depth)
        for mode, batches in ((False, ds), (True, hs), ("labels", hs)):
            for _ in range(2):
                runner.run_pipelined(batches, to_host=mode, consume=lambda k, r: None)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            a.record()
            runner.run_pipelined((batches[i % 6] for i in range(96)), to_host=mode, consume=lambda k, r: None)
            b.record()
            torch.cuda.synchronize()
            wall = (time.perf_counter() - t0) / 96 * 1e3
            print(json.dumps({"depth": depth, "host_slices": slices, "to_host": mode, "ms_per_batch": round(a.elapsed_time(b) / 96, 4),
                              "wall_ms_per_batch": round(wall, 4)}), flush=True)
# host-side cost of one submit (no device wait)
runner = GraphedSemSeg(net, depth=4)
runner.run_pipelined(ds, consume=lambda k, r: None)
torch.cuda.synchronize()
t0 = time.perf_counter()
ts = [runner.submit(ds[i % 6]) for i in range(4)]
t1 = time.perf_counter()
torch.cuda.synchronize()
print("host time per submit (us):", (t1 - t0) / 4 * 1e6)
