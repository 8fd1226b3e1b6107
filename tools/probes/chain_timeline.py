"""Per-phase clock64() timeline of the resident-weight chain kernel (CTA 0): fp1 + head, sa1 and sa2 shapes at C2 size.

    python tools/probes/chain_timeline.py [fp1|sa1|sa2]
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import _native as nv, ops  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "fp1"
B = 8


def mk(dims, last_relu):
    out = []
    for i, (ci, co) in enumerate(dims):
        out.append((torch.randn(co, ci, device=dev) * (2.0 / ci) ** 0.5, torch.randn(co, device=dev) * 0.1,
                    True if i + 1 < len(dims) else last_relu))
    return ops.PackedChain(out)


if which == "fp1":
    N, S = 24000, 1024
    chain = mk([(128, 128), (128, 128), (128, 128), (128, 19)], False)
    p2 = torch.randn(B, S, 128, device=dev)
    idx = torch.randint(0, S, (B, N, 3), device=dev)
    w = torch.rand(B, N, 3, device=dev)
    w = w / w.sum(-1, keepdim=True)
    run = lambda: ops.fp_mlp_tc(chain, None, p2, idx, w, ops.OUT_LOG_SOFTMAX, relu_in=True)
    nl = 4
else:
    N, S, D, dims = (24000, 1024, 1, [(4, 32), (32, 32), (32, 64)]) if which == "sa1" else (1024, 256, 64, [(67, 64), (64, 64), (64, 128)])
    chain = mk(dims, True)
    xyz = torch.rand(B, 3, N, device=dev).permute(0, 2, 1)
    feat = torch.randn(B, N, D, device=dev)
    q = xyz[:, :S].contiguous()
    idx = torch.randint(0, N, (B, S, 32), device=dev)
    run = lambda: ops.sa_mlp_max_tc(chain, xyz, feat, q, idx, False)
    nl = 3
dbg = torch.zeros(4 * 64 * 32, dtype=torch.int64, device=dev)
for _ in range(2):
    run()
with ops.options(mlp_debug=dbg.data_ptr()):   # pn_launch_opts.mlp_debug of every chain launch in this block
    run()
    torch.cuda.synchronize()
t = dbg.cpu().numpy().reshape(4, 64, 32)
t0 = t[t > 0].min()
names = ["start", "prod"] + sum([[f"L{l}.issue0", f"L{l}.issued", f"L{l}.ready", f"L{l}.epi"] for l in range(nl)], []) + ["done"]
for g in range(4):
    for r in range(3):
        row = t[g, r]
        if row[0] == 0:
            continue
        d = np.diff(row[:len(names)])
        print(f"group {g} round {r}: start={int(row[0] - t0)} total={int(row[len(names) - 1] - row[0])}  " +
              " ".join(f"{n}:{int(v)}" for n, v in zip(names[1:], d)))
