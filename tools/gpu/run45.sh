cd $GRAFT_REPO_ROOT
python - <<'PY' 2>&1 | grep -v Warn
import torch, time
from pointnet12_b200 import ops, synthetic as syn
dev = torch.device("cuda", 0)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev).permute(0, 2, 1)[:, :, :3]
torch.manual_seed(0)
st = torch.randint(0, 24000, (8,)).to(dev)
ref = ops.fps(x, 1024, st)
for cfg in ((8, 128, 2), (4, 256, 2), (3, 256, 2), (2, 256, 2), (2, 256, 3)):
    try:
        out = ops.fps(x, 1024, st, config=cfg)
        torch.cuda.synchronize()
        ts = []
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); ops.fps(x, 1024, st, config=cfg); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ctas = ops.fps_launch_info(8, 24000, 1024, cfg)
        print(cfg, "equal", torch.equal(out, ref), "ms", sorted(ts)[2], "launch", ctas)
    except Exception as e:
        print(cfg, "error", repr(e)[:200])
PY
timeout 900 python tools/pipeline_sweep.py --steps 96 --configs '{"depth": 6, "env": {}};{"depth": 6, "env": {"PN12_FPS1": "3,256,2"}};{"depth": 6, "env": {"PN12_FPS1": "2,256,2"}};{"depth": 8, "env": {"PN12_FPS1": "2,256,2"}};{"depth": 8, "env": {"PN12_FPS1": "3,256,2"}};{"depth": 6, "env": {}}' 2>&1 | grep depth
