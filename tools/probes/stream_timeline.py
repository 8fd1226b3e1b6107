"""(tag, clock) log of CTA 0 / thread 0 of the STREAMING chain kernel for the small levels of PointNet2SemSeg at C2 size.
tags: 1 tile start, 2 producer chunk done, 3 MMAs of a chunk issued, 4 accumulator ready, 5 epilogue of a pass done."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import _native as nv, ops  # noqa: E402

dev = torch.device("cuda", 0)
which = sys.argv[1] if len(sys.argv) > 1 else "fp4"
B = 8


def mk(dims, last_relu=True):
    out = []
    for i, (ci, co) in enumerate(dims):
        out.append((torch.randn(co, ci, device=dev) * (2.0 / ci) ** 0.5, torch.randn(co, device=dev) * 0.1,
                    True if i + 1 < len(dims) else last_relu))
    return ops.PackedChain(out)


cfg = {"fp4l1": (64, 16, 512, 256, [(768, 256)]), "fp2l1": (1024, 256, 64, 256, [(320, 256)]),
       "fp4": (64, 16, 512, 256, [(768, 256), (256, 256)]), "fp3": (256, 64, 256, 128, [(384, 256), (256, 256)]),
       "fp2": (1024, 256, 64, 256, [(320, 256), (256, 128)])}
if which in cfg:
    N, S, D1, D2, dims = cfg[which]
    chain = mk(dims)
    p1 = torch.randn(B, N, D1, device=dev)
    p2 = torch.randn(B, S, D2, device=dev)
    idx = torch.randint(0, S, (B, N, 3), device=dev)
    w = torch.rand(B, N, 3, device=dev)
    w = w / w.sum(-1, keepdim=True)
    run = lambda: ops.fp_mlp_tc(chain, p1, p2, idx, w, ops.OUT_ROWS)
else:
    N, S, D, dims = {"sa3": (256, 64, 128, [(131, 128), (128, 128), (128, 256)]),
                     "sa4": (64, 16, 256, [(259, 256), (256, 256), (256, 512)])}[which]
    chain = mk(dims)
    xyz = torch.rand(B, 3, N, device=dev).permute(0, 2, 1)
    feat = torch.randn(B, N, D, device=dev)
    q = xyz[:, :S].contiguous()
    idx = torch.randint(0, N, (B, S, 32), device=dev)
    run = lambda: ops.sa_mlp_max_tc(chain, xyz, feat, q, idx, False)
dbg = torch.zeros(1 + 2 * 512, dtype=torch.int64, device=dev)
for _ in range(2):
    run()
with ops.options(mlp_debug=dbg.data_ptr()):   # pn_launch_opts.mlp_debug of every chain launch in this block
    run()
    torch.cuda.synchronize()
t = dbg.cpu().numpy()
n = int(t[0])
names = {1: "tile", 2: "prod", 3: "issued", 4: "ready", 5: "epi", 6: "w", 7: "m"}
prev = t[2]
out = []
for i in range(n):
    tag, clk = int(t[1 + 2 * i]), int(t[2 + 2 * i])
    out.append(f"{names[tag]}+{clk - prev}")
    prev = clk
print(which, "total", int(t[2 * n] - t[2]), "cycles:", " ".join(out))
