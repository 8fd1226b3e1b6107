"""Generate tests/golden/*.npz by running the REFERENCE ITSELF (imported read-only from /root/reference).

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The reference is pure Python/PyTorch, so it cannot travel with the repo; its outputs on seeded
synthetic inputs do.  Inputs are rebuilt from seeds by pointnet12_b200.synthetic (a checksum of every
input is stored next to the outputs), FPS start indices come from torch.manual_seed(0) +
torch.randint exactly as pointnet_util.py:75 draws them, and index tensors are stored as int16/int32
to keep the fixtures small.  The shipped checkpoint (a data file, not source) is copied next to the
fixtures so that the GPU box can load it.
"""
import importlib.machinery
import importlib.util
import os
import shutil
import sys
import types

import numpy as np
import torch

REF = os.environ.get("PN_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "tests", "golden")
CKPT = "pointnet2-inview-0.55884-0001.pth"
sys.path.insert(0, ROOT)

from pointnet12_b200 import synthetic as syn  # noqa: E402


def import_reference():
    """Import /root/reference/model/*.py as the package `_pn12_ref` (the repo has its own `model`)."""
    pkg = types.ModuleType("_pn12_ref")
    pkg.__path__ = [os.path.join(REF, "model")]
    sys.modules["_pn12_ref"] = pkg
    mods = {}
    for name in ("pointnet_util", "pointnet2", "pointnet"):
        spec = importlib.util.spec_from_file_location(f"_pn12_ref.{name}", os.path.join(REF, "model", name + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        mods[name] = m
    return mods


def small(idx: torch.Tensor) -> np.ndarray:
    a = idx.numpy()
    return a.astype(np.int16 if a.max() < 32768 else np.int32)


def out_t(net, l1_xyz, l2_xyz, l1_f, p2):
    """fp2 output as a tensor (128 channels on the 1024 level-1 points): the input fp1 expects."""
    return net.fp2(l1_xyz, l2_xyz, l1_f, p2)


def save(name, **arrays):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB  keys={sorted(arrays)}")


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    ref = import_reference()
    U, P2, P1 = ref["pointnet_util"], ref["pointnet2"], ref["pointnet"]
    shutil.copyfile(os.path.join(REF, "checkpoints", CKPT), os.path.join(OUT, CKPT))
    ckpt = torch.load(os.path.join(OUT, CKPT), map_location="cpu")
    ckpt = {k[len("module."):]: v for k, v in ckpt.items()}

    # ---------------------------------------------------------------- L1 primitives on the C2 level chain
    pts = torch.from_numpy(syn.kitti_batch(2, 24000, config=2))            # [2,4,24000]
    xyz = pts[:, :3, :].permute(0, 2, 1)                                   # strided view like pointnet_util.py:184
    feat = pts[:, 3:, :].permute(0, 2, 1)
    out = {"input_sum": np.array(syn.checksum(pts.numpy()))}
    torch.manual_seed(0)
    cur = xyz
    for lvl, (npoint, radius) in enumerate([(1024, 0.1), (256, 0.2), (64, 0.4), (16, 0.8)], 1):
        st = torch.get_rng_state()
        start = torch.randint(0, cur.shape[1], (cur.shape[0],), dtype=torch.long)   # same draw as :75
        torch.set_rng_state(st)
        fps = U.farthest_point_sample(cur, npoint)
        assert torch.equal(fps[:, 0], start)
        new_xyz = U.index_points(cur, fps)
        ball = U.query_ball_point(radius, 32, cur, new_xyz)
        out[f"l{lvl}_start"] = start.numpy()
        out[f"l{lvl}_fps"] = small(fps)
        out[f"l{lvl}_ball"] = small(ball)
        out[f"l{lvl}_new_xyz"] = new_xyz.numpy()
        cur = new_xyz
    # square_distance: both argument orders on a modest slice (S, N >= 16)
    a, b = xyz[:, :64, :].contiguous(), xyz[:, 5000:5512, :]
    out["sqd_ab"] = U.square_distance(a, b).numpy()
    out["sqd_ba"] = U.square_distance(b, a).numpy()
    # sample_and_group at level 1 (grouped tensor, SSG order)
    torch.manual_seed(0)
    nx, npts = U.sample_and_group(1024, 0.1, 32, xyz, feat)
    out["sg_new_points_b0"] = npts[0, :64].numpy()                          # [64,32,4]
    save("primitives_c2", **out)

    # ---------------------------------------------------------------- L1 primitives, other shapes
    out = {}
    for n in (4096, 16384):
        p = torch.from_numpy(syn.kitti_batch(1, n, config=3))
        x = p[:, :3, :].permute(0, 2, 1)
        for npoint, radius in ((256, 0.2), (64, 0.4)):
            torch.manual_seed(n + npoint)
            fps = U.farthest_point_sample(x, npoint)
            ball = U.query_ball_point(radius, 32, x, U.index_points(x, fps))
            out[f"n{n}_p{npoint}_fps"] = small(fps)
            out[f"n{n}_p{npoint}_ball"] = small(ball)
    # ModelNet-shaped cloud, MSG radii / nsample (C4)
    m = torch.from_numpy(syn.modelnet_batch(2, 1024))
    mx = m.permute(0, 2, 1)
    torch.manual_seed(7)
    fps = U.farthest_point_sample(mx, 512)
    out["mn_fps"] = small(fps)
    for r, k in ((0.1, 16), (0.2, 32), (0.4, 128)):
        out[f"mn_ball_r{r}_k{k}"] = small(U.query_ball_point(r, k, mx, U.index_points(mx, fps)))
    save("primitives_misc", **out)

    # ---------------------------------------------------------------- L2 blocks with checkpoint weights
    net = P2.PointNet2SemSeg(19, feature_dims=1)
    net.load_state_dict(ckpt)
    net.eval()
    out = {}
    with torch.no_grad():
        p = torch.from_numpy(syn.kitti_batch(2, 4096, config=2))
        out["input_sum"] = np.array(syn.checksum(p.numpy()))
        torch.manual_seed(0)
        l1_xyz, l1_f = net.sa1(p[:, :3, :], p[:, 3:, :])
        l2_xyz, l2_f = net.sa2(l1_xyz, l1_f)
        out["sa1_xyz"], out["sa1_feat"] = l1_xyz.numpy(), l1_f.numpy()
        out["sa2_xyz"], out["sa2_feat"] = l2_xyz.numpy(), l2_f.numpy()
        # fp2 wants 256-channel coarse features (the output of fp3): use seeded noise of that width
        p2 = torch.from_numpy(np.random.default_rng(5).normal(0, 1, (2, 256, 256)).astype(np.float32))
        out["fp2_out"] = net.fp2(l1_xyz, l2_xyz, l1_f, p2).numpy()          # skip + interp, 320 -> 256 -> 128
        out["fp1_out_sub8"] = net.fp1(p[:, :3, :], l1_xyz, None, out_t(net, l1_xyz, l2_xyz, l1_f, p2))[:, :, ::8].numpy()
    save("blocks_ckpt", **out)

    # ---------------------------------------------------------------- PointNet2SemSeg, checkpoint
    out = {}
    with torch.no_grad():
        p = torch.from_numpy(syn.kitti_batch(2, 4096, config=2))
        torch.manual_seed(0)
        logp = net(p)
        out["n4096_logp"] = logp.numpy()                                    # [2,4096,19]
        p = torch.from_numpy(syn.kitti_batch(2, 24000, config=2))
        out["n24000_input_sum"] = np.array(syn.checksum(p.numpy()))
        torch.manual_seed(0)
        logp = net(p)
        out["n24000_logp_sub"] = logp[:, ::16, :].numpy()                   # every 16th point
        out["n24000_label"] = logp.argmax(-1).numpy().astype(np.uint8)
        top2 = logp.topk(2, dim=-1)[0]
        out["n24000_margin"] = (top2[..., 0] - top2[..., 1]).numpy().astype(np.float16)
    save("pointnet2_semseg_ckpt", **out)

    # ---------------------------------------------------------------- PointNetSeg, seeded weights
    net1 = P1.PointNetSeg(19, input_dims=4, feature_transform=True)
    sd = syn.random_state_dict({k: tuple(v.shape) for k, v in net1.state_dict().items()}, seed=1234)
    net1.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net1.eval()
    with torch.no_grad():
        p = torch.from_numpy(syn.kitti_batch(2, 2048, config=1))
        logp, tf = net1(p)
    save("pointnet_seg_seed1234", logp=logp.numpy(), trans_feat=tf.numpy(), input_sum=np.array(syn.checksum(p.numpy())))

    # ---------------------------------------------------------------- PointNet2ClsMsg, seeded weights
    net4 = P2.PointNet2ClsMsg()
    sd = syn.random_state_dict({k: tuple(v.shape) for k, v in net4.state_dict().items()}, seed=1234)
    net4.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    net4.eval()
    with torch.no_grad():
        m = torch.from_numpy(syn.modelnet_batch(4, 1024))
        torch.manual_seed(0)
        logp, l3 = net4(m)
    save("pointnet2_cls_msg_seed1234", logp=logp.numpy(), l3_points=l3.numpy())

    # ---------------------------------------------------------------- remaining heads (f-4), seeded weights
    out = {}
    with torch.no_grad():
        n = P2.PointNet2ClsSsg()
        sd = syn.random_state_dict({k: tuple(v.shape) for k, v in n.state_dict().items()}, seed=77)
        n.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        n.eval()
        torch.manual_seed(0)
        out["cls_ssg_logp"] = n(torch.from_numpy(syn.modelnet_batch(2, 1024))).numpy()
        n = P2.PointNet2PartSegSsg(50)
        sd = syn.random_state_dict({k: tuple(v.shape) for k, v in n.state_dict().items()}, seed=78)
        n.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        n.eval()
        torch.manual_seed(0)
        x, f = n(torch.from_numpy(syn.modelnet_batch(2, 1024)))
        out["partseg_ssg_logp"], out["partseg_ssg_feat"] = x.numpy(), f.numpy()
        n = P1.PointNetCls(k=40, feature_transform=True)
        sd = syn.random_state_dict({k: tuple(v.shape) for k, v in n.state_dict().items()}, seed=79)
        n.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
        n.eval()
        x, tf = n(torch.from_numpy(syn.modelnet_batch(2, 1024)))
        out["pointnet_cls_logp"], out["pointnet_cls_tf"] = x.numpy(), tf.numpy()
    save("other_heads_seeded", **out)


if __name__ == "__main__":
    main()
