"""Times pn_train_gemm_bf16x3 on one layer shape (PN12_GEMM_DEBUG selects which part of the kernel is switched off)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops  # noqa: E402
dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for rows, cin, cout in [(64000, 128, 128), (262144, 32, 64)]:
    x = torch.randn(rows, cin, device=dev); w = torch.randn(cout, cin, device=dev); b = torch.randn(cout, device=dev)
    g = torch.cuda.CUDAGraph()
    acc = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    ops.train_gemm(x, w, b, stats_acc=acc); torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(10):
            ops.train_gemm(x, w, b, stats_acc=acc)
    ts = []
    for _ in range(5):
        flush.fill_(1)
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); e.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(e) * 100)
    print(os.environ.get("PN12_GEMM_DEBUG", "0"), rows, cin, cout, f"{sorted(ts)[2]:.1f} us per call (10 calls in a graph)")
