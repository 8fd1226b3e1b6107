"""One C2 forward after warm-up, for ncu:  ncu ... python tools/profile_forward.py [warm] """
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pointnet12_b200 import synthetic as syn  # noqa: E402
from pointnet12_b200.model.utils import load_pointnet  # noqa: E402

warm = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device("cuda", 0)
net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
with torch.no_grad():
    for i in range(warm + 1):
        torch.manual_seed(0)
        if i == warm:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()          # ncu --profile-from-start off: only the last forward is profiled
        net(x)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
