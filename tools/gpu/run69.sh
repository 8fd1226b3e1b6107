cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fps or preprocess or empty_scan or chamfer or pipelined" 2>&1 | tail -3
python - <<'PY' 2>&1 | grep -v Warn
import torch
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
for B, N in ((1, 120000), (8, 120000), (1, 98304), (1, 65536), (8, 65536)):
    x = torch.from_numpy(syn.kitti_batch(B, N, config=3)).to(dev).permute(0, 2, 1)[:, :, :3]
    torch.manual_seed(0)
    st = torch.randint(0, N, (B,)).to(dev)
    ref = None
    for cfg in ((0, 0, 0), (16, 512, 0), (16, 256, 2), (8, 256, 2)):
        try:
            out = ops.fps(x, 1024, st, config=cfg)
            torch.cuda.synchronize()
            if ref is None: ref = out
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.fps(x, 1024, st, config=cfg); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            ms = sorted(ts)[2]
            print(B, N, cfg, "equal", torch.equal(out, ref), "ms", round(ms, 4), "TB/s", round(B * 1024 * N * 16 / ms / 1e9, 2), ops.fps_launch_info(B, N, 1024, cfg))
        except Exception as e:
            print(B, N, cfg, "error", repr(e)[:120])
PY
