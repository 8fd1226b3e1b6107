cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k "fps or golden_chain or pipelined" 2>&1 | tail -2
python - <<'PY' 2>&1 | grep -v Warn
import torch
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev).permute(0, 2, 1)[:, :, :3]
torch.manual_seed(0)
st = torch.randint(0, 24000, (8,)).to(dev)
for cfg in ((8, 128, 2), (4, 256, 2), (3, 256, 2), (2, 256, 2)):
    out = ops.fps(x, 1024, st, config=cfg)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); ops.fps(x, 1024, st, config=cfg); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(cfg, "ms", sorted(ts)[2])
PY
timeout 900 python tools/pipeline_sweep.py --steps 192 --configs '{"depth": 10, "env": {}};{"depth": 10, "env": {}}' 2>&1 | grep "depth"
