cd $GRAFT_REPO_ROOT
timeout 300 python - <<'PY' 2>&1 | grep -v Warn
import torch, numpy as np
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
for B in (8, 3, 2, 1):
    x = torch.from_numpy(syn.kitti_batch(B, 24000, config=2)).to(dev).permute(0, 2, 1)[:, :, :3]
    torch.manual_seed(0)
    st = torch.randint(0, 24000, (B,)).to(dev)
    ref = ops.fps(x, 1024, st)
    for cfg in ((2, 256, 2), (4, 256, 4), (2, 256, 4), (3, 256, 4), (8, 256, 4)):
        try:
            out = ops.fps(x, 1024, st, config=cfg)
            torch.cuda.synchronize()
            ts = []
            for _ in range(5):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); ops.fps(x, 1024, st, config=cfg); b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            print("B", B, cfg, "equal", torch.equal(out, ref), "ms", round(sorted(ts)[2], 4), "launch", ops.fps_launch_info(B, 24000, 1024, cfg))
        except Exception as e:
            print("B", B, cfg, "error", repr(e)[:160])
PY
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {"PN12_FPS1": "4,256,4"}};{"depth": 12, "env": {"PN12_FPS1": "4,256,4"}};{"depth": 16, "env": {"PN12_FPS1": "4,256,4"}};{"depth": 10, "env": {"PN12_FPS1": "3,256,4"}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
