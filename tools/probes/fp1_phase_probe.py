"""Timing probes of the fp1 + head chain at C2 (results invalid in probe modes): full kernel, producer only, chain only,
one warp group only -- which phase bounds the tile rate?   python tools/probes/fp1_phase_probe.py"""
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402

dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
n = net.module
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
pm = x.permute(0, 2, 1)
x0 = pm[:, :, :3]
torch.manual_seed(0)
with torch.no_grad():
    fps1 = ops.fps(x0, 1024, torch.randint(0, 24000, (8,)).to(dev))
    x1 = ops.index_points(x0, fps1)
    idx, w = ops.three_nn(x0, x1, method="scan")
    z = torch.randn(8, 1024, 128, device=dev)
    head = (n._head, [n.conv1, n.conv2], [n.bn1, None], [True, False], ops.OUT_LOG_SOFTMAX)
    split = n._head.chain_folded_first(list(n.fp1.mlp_convs) + [n.conv1, n.conv2], list(n.fp1.mlp_bns) + [n.bn1, None], [True] * 4 + [False])
    rest = split[1]

    def run():
        return ops.fp_mlp_tc(rest, None, z, idx, w, ops.OUT_LOG_SOFTMAX, relu_in=True)

    for name, eng in (("full", 0), ("full, generic producer", 256), ("full, no MMA lock", 512), ("full, generic producer, no lock (round 1)", 768),
                      ("producer only", 32), ("producer only, generic", 32 + 256), ("chain only (no producer)", 128),
                      ("chain only, no lock", 128 + 512), ("one group: full", 64), ("one group: producer only", 96), ("one group: chain only", 192)):
        ops._DEFAULTS["mlp_engine"] = eng
        ts = []
        for i in range(6):
            flush.fill_(i)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        print(f"{name:44s} {statistics.median(ts[1:]):8.1f} us")
    ops._DEFAULTS["mlp_engine"] = 0
