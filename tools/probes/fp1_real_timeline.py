"""clock64 timeline of CTA 0 of the LAST resident-chain launch (fp1 + head) of a real C2 forward (checkpoint weights,
synthetic KITTI-shaped clouds): what the phases cost with real neighbour indices, a cold L2 and the real output stream."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops, synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402

dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
dbg = torch.zeros(4 * 64 * 32, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(3):
        net(x)
    flush.fill_(1)
    with ops.options(mlp_debug=dbg.data_ptr()):
        net(x)
        torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(4, 64, 32)
t0 = t[t > 0].min()
names = ["start", "prod"] + sum([[f"L{l}.issue0", f"L{l}.issued", f"L{l}.ready", f"L{l}.epi"] for l in range(4)], []) + ["done"]
for g in range(2):
    for r in range(7):
        row = t[g, r]
        if row[0] == 0:
            continue
        d = np.diff(row[:len(names)])
        print(f"group {g} round {r}: start={int(row[0] - t0)} total={int(row[len(names) - 1] - row[0])}  " +
              " ".join(f"{n}:{int(v)}" for n, v in zip(names[1:], d)) + "  | producer: " + " ".join(str(int(v - row[0])) for v in row[22:28] if v > 0))
