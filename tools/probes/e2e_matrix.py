"""Which half of the end-to-end path costs throughput: input from the host or output to the host?  ms per batch, depth 6."""
import json
import os
import sys

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
runner = GraphedSemSeg(net, depth=int(os.environ.get("DEPTH", "6")))
for rep in range(2):
    for name_in, batches in (("device", ds), ("host", hs)):
        for mode in (False, True, "labels"):
            for _ in range(2):
                runner.run_pipelined(batches, to_host=mode, consume=lambda k, r: None)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            runner.run_pipelined((batches[i % 6] for i in range(96)), to_host=mode, consume=lambda k, r: None)
            b.record()
            torch.cuda.synchronize()
            print(json.dumps({"input": name_in, "to_host": mode, "ms_per_batch": round(a.elapsed_time(b) / 96, 4)}), flush=True)
