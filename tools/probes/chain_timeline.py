"""Per-phase clock64() timeline of the resident-weight chain kernel (CTA 0) for the fp1 + head chain at C2 size."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import _native as nv, ops  # noqa: E402

dev = torch.device("cuda", 0)
rng = np.random.default_rng(0)
B, N, S = 8, 24000, 1024
layers = []
for ci, co in [(128, 128), (128, 128), (128, 128), (128, 19)]:
    layers.append((torch.randn(co, ci, device=dev) * (2.0 / ci) ** 0.5, torch.randn(co, device=dev) * 0.1, co != 19))
chain = ops.PackedChain(layers)
p2 = torch.randn(B, S, 128, device=dev)
idx = torch.randint(0, S, (B, N, 3), device=dev)
w = torch.rand(B, N, 3, device=dev)
w = w / w.sum(-1, keepdim=True)
dbg = torch.zeros(4 * 64 * 32, dtype=torch.int64, device=dev)
for _ in range(2):
    ops.fp_mlp_tc(chain, None, p2, idx, w, ops.OUT_LOG_SOFTMAX, relu_in=True)
nv.call("pn_mlp_set_debug", dbg.data_ptr())
ops.fp_mlp_tc(chain, None, p2, idx, w, ops.OUT_LOG_SOFTMAX, relu_in=True)
torch.cuda.synchronize()
nv.call("pn_mlp_set_debug", None)
t = dbg.cpu().numpy().reshape(4, 64, 32)
t0 = t[t > 0].min()
names = ["start", "prod"] + sum([[f"L{l}.issue0", f"L{l}.issued", f"L{l}.ready", f"L{l}.epi"] for l in range(4)], []) + ["done"]
for g in range(2):
    for r in range(8):
        row = t[g, r]
        if row[0] == 0:
            continue
        rel = row[:len(names)] - t0
        print(f"group {g} round {r}: " + " ".join(f"{n}={int(v)}" for n, v in zip(names, rel)))
        d = np.diff(row[:len(names)])
        print("    deltas: " + " ".join(f"{n}:{int(v)}" for n, v in zip(names[1:], d)))
