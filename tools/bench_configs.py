"""Secondary configurations of BASELINE.json (not the bench.py line): CUDA-event timings of
  C1  PointNetSeg(19, input_dims=4, feature_transform) forward, B=1, N=24000 (seeded random init: the reference ships no
      pointnet-inview checkpoint)
  C4  PointNet2ClsMsg forward, B=32, N=1024 ModelNet40-shaped clouds, in bf16x3 (fp32 parity) and bf16
Inputs resident on the device, 3 warm-ups, median of 20, L2 flushed between iterations.  C1 is also timed as one CUDA-graph
replay (runtime.GraphedModule: PointNetSeg draws nothing on the host), and the host link of the box is probed (pinned
device-to-host copy of 14.6 MB, the size of one C2 output) to put the end-to-end figures of bench.py into context."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402
from pointnet12_b200.model.pointnet import PointNetSeg  # noqa: E402
from pointnet12_b200.model.pointnet2 import PointNet2ClsMsg  # noqa: E402
from pointnet12_b200.runtime import GraphedModule  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def seeded(net, seed):
    sd = syn.random_state_dict({k: tuple(v.shape) for k, v in net.state_dict().items()}, seed)
    net.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    return net.to(dev).eval()


def median_ms(fn, iters=20, warmup=3):
    for _ in range(warmup):
        fn()
    ts = []
    for i in range(iters):
        flush.fill_(i & 0xFF)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


with torch.no_grad():
    net = seeded(PointNetSeg(19, input_dims=4, feature_transform=True), 1234)
    x = torch.from_numpy(syn.kitti_batch(1, 24000, config=1)).to(dev)
    for mode in ("bf16x3", "bf16", "fp32"):
        ops.set_mlp_mode(mode)
        ms = median_ms(lambda: net(x))
        print(json.dumps({"config": "C1 PointNetSeg B=1 N=24000", "precision": mode, "ms": round(ms, 4),
                          "points_per_s": round(24000 / ms * 1e3)}), flush=True)
    ops.set_mlp_mode("bf16x3")
    runner = GraphedModule(net)
    ms = median_ms(lambda: runner(x))
    print(json.dumps({"config": "C1 PointNetSeg B=1 N=24000, one CUDA-graph replay", "precision": "bf16x3", "ms": round(ms, 4),
                      "points_per_s": round(24000 / ms * 1e3)}), flush=True)
    net = seeded(PointNet2ClsMsg(), 1234)
    x = torch.from_numpy(syn.modelnet_batch(32, 1024)).to(dev)
    for mode in ("bf16x3", "bf16", "fp32"):
        ops.set_mlp_mode(mode)
        ms = median_ms(lambda: net(x))
        print(json.dumps({"config": "C4 PointNet2ClsMsg B=32 N=1024", "precision": mode, "ms": round(ms, 4),
                          "clouds_per_s": round(32 / ms * 1e3), "tflops_useful": round(250.6 / ms, 1)}), flush=True)
        if mode != "fp32":
            runner = GraphedModule(net)
            ms = median_ms(lambda: runner(x))
            print(json.dumps({"config": "C4 PointNet2ClsMsg B=32 N=1024, one CUDA-graph replay", "precision": mode, "ms": round(ms, 4),
                              "clouds_per_s": round(32 / ms * 1e3), "tflops_useful": round(250.6 / ms, 1)}), flush=True)
ops.set_mlp_mode("bf16x3")

# host link: pinned device-to-host / host-to-device copies of one C2 output / input
out_d = torch.empty((8, 24000, 19), dtype=torch.float32, device=dev)
out_h = torch.empty((8, 24000, 19), dtype=torch.float32).pin_memory()
in_h = torch.empty((8, 4, 24000), dtype=torch.float32).pin_memory()
in_d = torch.empty((8, 4, 24000), dtype=torch.float32, device=dev)
for name, fn, nbytes in (("d2h 14.6 MB", lambda: out_h.copy_(out_d, non_blocking=True), out_d.numel() * 4),
                         ("h2d 3.1 MB", lambda: in_d.copy_(in_h, non_blocking=True), in_h.numel() * 4)):
    ms = median_ms(fn)
    print(json.dumps({"config": "host link, pinned " + name, "ms": round(ms, 4), "GB_per_s": round(nbytes / ms / 1e6, 1)}), flush=True)
