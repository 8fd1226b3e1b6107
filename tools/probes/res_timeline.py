"""clock64 timeline of CTA 0 of one resident-weight chain launch of a real C2 forward (sa1, sa2 or fp1 + head), run alone
on the real intermediate tensors with a cold L2.   python tools/probes/res_timeline.py sa1|sa2|fp1"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "sa1"
if len(sys.argv) > 2:
    ops._DEFAULTS["mlp_engine"] = int(sys.argv[2])       # e.g. 512 = no MMA issue lock, 256 = generic producers
dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
n = net.module
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
pm = x.permute(0, 2, 1)
x0, f0 = pm[:, :, :3], pm[:, :, 3:]
torch.manual_seed(0)
st = [torch.randint(0, m, (8,)).to(dev) for m in (24000, 1024, 256, 64)]
dbg = torch.zeros(4 * 64 * 32, dtype=torch.int64, device=dev)
with torch.no_grad():
    x1, b1 = n.sa1.geometry(x0, st[0])
    f1 = n.sa1.features(x0, f0, x1, b1)
    x2, b2 = n.sa2.geometry(x1, st[1])
    if which == "sa1":
        run, nl, groups = (lambda: n.sa1.features(x0, f0, x1, b1)), 3, 4
    elif which == "sa2":
        run, nl, groups = (lambda: n.sa2.features(x1, f1, x2, b2)), 3, 2
    else:
        idx, w = ops.three_nn(x0, x1, method="scan")
        z = torch.randn(8, 1024, 128, device=dev)
        rest = n._head.chain_folded_first(list(n.fp1.mlp_convs) + [n.conv1, n.conv2], list(n.fp1.mlp_bns) + [n.bn1, None], [True] * 4 + [False])[1]
        run, nl, groups = (lambda: ops.fp_mlp_tc(rest, None, z, idx, w, ops.OUT_LOG_SOFTMAX, relu_in=True)), 4, 2
    for _ in range(2):
        run()
    flush.fill_(1)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ops.options(mlp_debug=dbg.data_ptr()):
        a.record()
        run()
        b.record()
        torch.cuda.synchronize()
print(f"{which}: {a.elapsed_time(b) * 1e3:.1f} us")
t = dbg.cpu().numpy().reshape(4, 64, 32)
names = ["start", "prod"] + sum([[f"L{l}.issue0", f"L{l}.issued", f"L{l}.ready", f"L{l}.epi"] for l in range(nl)], []) + ["done"]
t0 = t[t > 0].min()
for g in range(groups):
    for r in (0, 1, 2, 3, 4):
        row = t[g, r]
        if row[0] == 0:
            continue
        d = np.diff(row[:len(names)])
        print(f"group {g} round {r}: start={int(row[0] - t0)} total={int(row[len(names) - 1] - row[0])}  " +
              " ".join(f"{nm}:{int(v)}" for nm, v in zip(names[1:], d)))
