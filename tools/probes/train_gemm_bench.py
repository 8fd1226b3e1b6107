"""Times pn_train_gemm_bf16x3 variants against the single-layer chain GEMM and the HBM traffic bound."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from pointnet12_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


for rows, cin, cout in [(64000, 128, 128), (262144, 32, 64), (262144, 4, 32), (65536, 64, 128), (16384, 128, 256), (64000, 128, 19)]:
    x = torch.randn(rows, cin, device=dev)
    w = torch.randn(cout, cin, device=dev)
    b = torch.randn(cout, device=dev)
    st = ops.BatchStats()
    st.scale = torch.rand(cin, device=dev) + 0.5
    st.shift = torch.randn(cin, device=dev)
    acc = torch.zeros(2, cout, dtype=torch.float64, device=dev)
    chain = ops.PackedChain([(w, b, False)])
    mb = rows * (cin + cout) * 4 / 1e6
    print(f"{rows}x{cin}->{cout}: traffic {mb:.0f} MB = {mb / 6.5e3 * 1e3:.1f} us at 6.5 TB/s | "
          f"plain {timeit(lambda: ops.train_gemm(x, w, b)):.1f} us, +transform {timeit(lambda: ops.train_gemm(x, w, b, in_stats=st)):.1f}, "
          f"+stats {timeit(lambda: ops.train_gemm(x, w, b, stats_acc=acc)):.1f}, both {timeit(lambda: ops.train_gemm(x, w, b, in_stats=st, stats_acc=acc)):.1f} | "
          f"chain {timeit(lambda: ops.mlp_rows_tc(chain, x)):.1f} us | linear fp32 {timeit(lambda: ops.linear(x, w, b, relu=False)):.1f} us")
