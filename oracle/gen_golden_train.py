"""Generate tests/golden/train_step_ckpt.npz by running the REFERENCE'S OWN training iteration (pcdseg.py:166-186)
on the CPU in the build container (the reference is imported read-only from /root/reference):

    python oracle/gen_golden_train.py

model = PointNet2SemSeg(19, feature_dims=1) with the shipped checkpoint, model.train(), torch.manual_seed(0),
logits = model(points); loss = nn.CrossEntropyLoss()(logits.transpose(2, 1), target); loss.backward().
Stored: the loss, the log-probabilities, the dropout keep-mask the reference drew (captured with a forward hook on
drop1; where the dropout input is 0 the mask is unobservable and irrelevant and is stored as 1), every parameter gradient
(complete for tensors up to 20000 elements, every 9th element + sum + L2 norm for the larger ones), and the BatchNorm
buffers after the forward.  Inputs are rebuilt from seeds (pointnet12_b200.synthetic, config 5), labels from
numpy default_rng(5000).
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from gen_golden import CKPT, OUT, REF, import_reference  # noqa: E402
from pointnet12_b200 import synthetic as syn  # noqa: E402
from oracle import oracle as orc  # noqa: E402

B, N, CLASSES = 2, 2048, 19
FULL_MAX, STRIDE = 20000, 9


def seeded_block(ctor, seed: int):
    """A block with seeded default init, then BatchNorm gamma ~ U(0.5, 1.5), beta ~ N(0, 0.1) from the same generator.
    tests/ rebuild OUR block the same way; main() asserts that both constructions give identical parameters."""
    torch.manual_seed(seed)
    blk = ctor()
    with torch.no_grad():
        for bn in blk.mlp_bns:
            bn.weight.copy_(torch.rand(bn.weight.shape) + 0.5)
            bn.bias.copy_(torch.randn(bn.bias.shape) * 0.1)
    return blk.train()


def block_inputs():
    """Seeded inputs of the two block fixtures: 1024 points of a synthetic cloud, 64-channel features, a coarse level
    of 256 points (an FPS sample of the cloud) with 256-channel features, and the upstream gradients."""
    rng = np.random.default_rng(6000)
    xyz = np.ascontiguousarray(syn.kitti_batch(2, 1024, config=6)[:, :3, :])                 # [2,3,1024]
    f1 = rng.standard_normal((2, 64, 1024)).astype(np.float32)
    # coarse level: 256 DISTINCT points (an FPS sample, like the network's own levels; duplicates would tie in the 3-NN sort)
    pm = np.ascontiguousarray(xyz.transpose(0, 2, 1))
    xyz2 = np.ascontiguousarray(orc.index_points(pm, orc.farthest_point_sample(pm, 256, [0, 0])).transpose(0, 2, 1))
    f2 = rng.standard_normal((2, 256, 256)).astype(np.float32)
    g_sa = rng.standard_normal((2, 128, 256)).astype(np.float32)
    g_fp = rng.standard_normal((2, 128, 1024)).astype(np.float32)
    return xyz, f1, xyz2, f2, g_sa, g_fp


def blocks(ref):
    """PointNetSetAbstraction(256, 0.2, 32, 67, [64,64,128]) and PointNetFeaturePropagation(320, [256,128]) of the
    reference in train mode: outputs, input gradients and parameter gradients for seeded inputs."""
    from pointnet12_b200.model import pointnet_util as ours

    pu = ref["pointnet_util"]
    xyz, f1, xyz2, f2, g_sa, g_fp = block_inputs()
    arrays = {}
    sa = seeded_block(lambda: pu.PointNetSetAbstraction(256, 0.2, 32, 64 + 3, [64, 64, 128], False), 4321)
    mine = seeded_block(lambda: ours.PointNetSetAbstraction(256, 0.2, 32, 64 + 3, [64, 64, 128], False), 4321)
    assert all(torch.equal(a, b) for a, b in zip(sa.state_dict().values(), mine.state_dict().values()))
    pts = torch.from_numpy(f1).requires_grad_(True)
    torch.manual_seed(7)
    start = torch.randint(0, 1024, (2,), dtype=torch.long).numpy()
    torch.manual_seed(7)
    new_xyz, new_points = sa(torch.from_numpy(xyz), pts)
    new_points.backward(torch.from_numpy(g_sa))
    arrays.update({"sa.start": start.astype(np.int32), "sa.out": new_points.detach().numpy(), "sa.dpoints": pts.grad.numpy()})
    for n, p in sa.named_parameters():
        arrays["sa.grad." + n] = p.grad.numpy()
    for n, b in sa.named_buffers():
        arrays["sa.buffer." + n] = b.numpy()
    fp = seeded_block(lambda: pu.PointNetFeaturePropagation(320, [256, 128]), 4322)
    mine = seeded_block(lambda: ours.PointNetFeaturePropagation(320, [256, 128]), 4322)
    assert all(torch.equal(a, b) for a, b in zip(fp.state_dict().values(), mine.state_dict().values()))
    p1 = torch.from_numpy(f1).requires_grad_(True)
    p2 = torch.from_numpy(f2).requires_grad_(True)
    out = fp(torch.from_numpy(xyz), torch.from_numpy(xyz2), p1, p2)
    out.backward(torch.from_numpy(g_fp))
    arrays.update({"fp.out": out.detach().numpy(), "fp.dpoints1": p1.grad.numpy(), "fp.dpoints2": p2.grad.numpy()})
    for n, p in fp.named_parameters():
        arrays["fp.grad." + n] = p.grad.numpy()
    path = os.path.join(OUT, "train_blocks_seeded.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def cls_nets(ref):
    """PointNet2ClsSsg / PointNet2ClsMsg (pointnet2.py:7-73) in train mode: forward, F.nll_loss, backward on 4 ModelNet40-shaped
    clouds; seeded default init (the test rebuilds OUR net under the same seed; asserted identical here).  The two dropout
    masks are captured with forward hooks (1 where the dropout input is 0)."""
    from pointnet12_b200.model import pointnet2 as ours

    arrays = {}
    xyz = syn.modelnet_batch(4, 1024, seed=4100)
    target = np.random.default_rng(4100).integers(0, 40, size=(4,)).astype(np.int64)
    arrays["target"] = target
    for tag, cls_name, nstarts in (("ssg", "PointNet2ClsSsg", (1024, 512)), ("msg", "PointNet2ClsMsg", (1024, 512))):
        torch.manual_seed(4242)
        net = getattr(ref["pointnet2"], cls_name)().train()
        torch.manual_seed(4242)
        mine = getattr(ours, cls_name)()
        assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), mine.state_dict().values())), cls_name
        seen = {}
        net.drop1.register_forward_hook(lambda m, i, o: seen.update(x1=i[0].detach().clone(), y1=o.detach().clone()))
        net.drop2.register_forward_hook(lambda m, i, o: seen.update(x2=i[0].detach().clone(), y2=o.detach().clone()))
        torch.manual_seed(9)
        starts = [torch.randint(0, n, (4,), dtype=torch.long).numpy() for n in nstarts]
        torch.manual_seed(9)
        out = net(torch.from_numpy(xyz))
        logp = out[0] if isinstance(out, tuple) else out
        loss = torch.nn.functional.nll_loss(logp, torch.from_numpy(target))
        net.zero_grad()
        loss.backward()
        arrays[f"{tag}.starts"] = np.stack(starts).astype(np.int32)
        arrays[f"{tag}.logp"] = logp.detach().numpy()
        arrays[f"{tag}.loss"] = np.float64(loss.item())
        for k in (1, 2):
            x, y = seen[f"x{k}"], seen[f"y{k}"]
            arrays[f"{tag}.keep{k}"] = torch.where(x != 0, y != 0, torch.ones_like(x, dtype=torch.bool)).numpy().astype(np.uint8)
        for name, p in net.named_parameters():
            g = p.grad.detach().numpy().reshape(-1)
            arrays[f"{tag}.grad.{name}"] = (g if g.size <= FULL_MAX else g[::STRIDE]).astype(np.float32)
        for name, b in net.named_buffers():
            if not name.endswith("num_batches_tracked"):
                arrays[f"{tag}.buffer.{name}"] = b.detach().numpy()
        print(f"{cls_name}: loss {loss.item():.5f}")
    # PointNet2PartSegSsg (pointnet2.py:75-104): two outputs (log-probs, features before dropout), both in the loss
    torch.manual_seed(4343)
    net = ref["pointnet2"].PointNet2PartSegSsg(50).train()
    torch.manual_seed(4343)
    mine = ours.PointNet2PartSegSsg(50)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), mine.state_dict().values()))
    xyz2 = syn.modelnet_batch(2, 1024, seed=4200)
    tgt2 = np.random.default_rng(4200).integers(0, 50, size=(2, 1024)).astype(np.int64)
    seen = {}
    net.drop1.register_forward_hook(lambda m, i, o: seen.update(x=i[0].detach().clone(), y=o.detach().clone()))
    torch.manual_seed(9)
    starts = [torch.randint(0, n, (2,), dtype=torch.long).numpy() for n in (1024, 512)]
    torch.manual_seed(9)
    logp, feat = net(torch.from_numpy(xyz2))
    loss = torch.nn.functional.nll_loss(logp.reshape(-1, 50), torch.from_numpy(tgt2).reshape(-1)) + 1e-3 * feat.pow(2).mean()
    net.zero_grad()
    loss.backward()
    keep = torch.where(seen["x"] != 0, seen["y"] != 0, torch.ones_like(seen["x"], dtype=torch.bool))      # [B,128,N]
    arrays.update({"part.target": tgt2.astype(np.int8), "part.logp": logp.detach().numpy().astype(np.float32),
                   "part.feat_sub": feat.detach().numpy()[:, :, ::8], "part.loss": np.float64(loss.item()),
                   "part.keep_bits": np.packbits(keep.permute(0, 2, 1).reshape(2 * 1024, 128).numpy(), axis=1)})
    for name, p in net.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        arrays[f"part.grad.{name}"] = (g if g.size <= FULL_MAX else g[::STRIDE]).astype(np.float32)
    print(f"PointNet2PartSegSsg: loss {loss.item():.5f}")
    path = os.path.join(OUT, "train_cls_seeded.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def pointnet_seg(ref):
    """PointNetSeg(19, input_dims=4, feature_transform=True) -- the default model of the reference's training driver
    (pcdseg.py:49-52, :128) -- one iteration as pcdseg.py:166-186 runs it: CrossEntropyLoss on the transposed log-probs plus
    0.001 * feature_transform_reguliarzer(trans_feat).  Gradients: every 31st element of the large tensors."""
    from pointnet12_b200.model import pointnet as ours

    rp = ref["pointnet"]
    torch.manual_seed(4444)
    net = rp.PointNetSeg(19, input_dims=4, feature_transform=True).train()
    torch.manual_seed(4444)
    mine = ours.PointNetSeg(19, input_dims=4, feature_transform=True)
    assert all(torch.equal(a, b) for a, b in zip(net.state_dict().values(), mine.state_dict().values()))
    # 8 clouds: the STNs' fully connected layers normalise over the BATCH -- with two clouds a channel whose two values
    # nearly coincide turns round-off into +-1 after BatchNorm
    pts = syn.kitti_batch(8, 512, config=7)
    target = np.random.default_rng(7000).integers(0, 19, size=(8, 512)).astype(np.int64)
    logits, trans_feat = net(torch.from_numpy(pts))
    loss = torch.nn.CrossEntropyLoss()(logits.transpose(2, 1), torch.from_numpy(target))
    loss = loss + rp.feature_transform_reguliarzer(trans_feat) * 0.001
    net.zero_grad()
    loss.backward()
    arrays = {"target": target.astype(np.int8), "logp": logits.detach().numpy(), "trans_feat": trans_feat.detach().numpy(),
              "loss": np.float64(loss.item())}
    for name, p in net.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        arrays["grad." + name] = (g if g.size <= 4096 else g[::31]).astype(np.float32)
    for name, b in net.named_buffers():
        if not name.endswith("num_batches_tracked"):
            arrays["buffer." + name] = b.detach().numpy()
    # PointNetCls(40, feature_transform=True) (pointnet.py:133-151): dropout sits between fc2 and bn2
    torch.manual_seed(4545)
    cnet = rp.PointNetCls(40, feature_transform=True).train()
    torch.manual_seed(4545)
    cmine = ours.PointNetCls(40, feature_transform=True)
    assert all(torch.equal(a, b) for a, b in zip(cnet.state_dict().values(), cmine.state_dict().values()))
    cx = syn.modelnet_batch(8, 512, seed=4300)
    ctgt = np.random.default_rng(4300).integers(0, 40, size=(8,)).astype(np.int64)
    seen = {}
    cnet.dropout.register_forward_hook(lambda m, i, o: seen.update(x=i[0].detach().clone(), y=o.detach().clone()))
    torch.manual_seed(3)
    clogp, ctf = cnet(torch.from_numpy(cx))
    closs = torch.nn.functional.nll_loss(clogp, torch.from_numpy(ctgt)) + rp.feature_transform_reguliarzer(ctf) * 0.001
    cnet.zero_grad()
    closs.backward()
    arrays.update({"cls.target": ctgt, "cls.logp": clogp.detach().numpy(), "cls.loss": np.float64(closs.item()),
                   "cls.keep": torch.where(seen["x"] != 0, seen["y"] != 0, torch.ones_like(seen["x"], dtype=torch.bool)).numpy().astype(np.uint8)})
    for name, p in cnet.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        arrays["cls.grad." + name] = (g if g.size <= 4096 else g[::31]).astype(np.float32)
    print(f"PointNetCls: loss {closs.item():.5f}")
    # PointNetDenseCls(16, 50) (pointnet.py:153-228): two heads sharing the encoder, 4944-channel concat
    torch.manual_seed(4646)
    dnet = rp.PointNetDenseCls(16, 50).train()
    torch.manual_seed(4646)
    dmine = ours.PointNetDenseCls(16, 50)
    assert all(torch.equal(a, b) for a, b in zip(dnet.state_dict().values(), dmine.state_dict().values()))
    dx = syn.modelnet_batch(8, 256, seed=4400)
    rng = np.random.default_rng(4400)
    dlabel = np.eye(16, dtype=np.float32)[rng.integers(0, 16, 8)]
    dcls, dseg = rng.integers(0, 16, size=(8,)).astype(np.int64), rng.integers(0, 50, size=(8, 256)).astype(np.int64)
    seen = {}
    dnet.dropout.register_forward_hook(lambda m, i, o: seen.update(x=i[0].detach().clone(), y=o.detach().clone()))
    torch.manual_seed(3)
    net1, net2, dtf = dnet(torch.from_numpy(dx), torch.from_numpy(dlabel))
    dloss = (torch.nn.functional.cross_entropy(net1, torch.from_numpy(dcls))
             + torch.nn.functional.nll_loss(net2.reshape(-1, 50), torch.from_numpy(dseg).reshape(-1))
             + rp.feature_transform_reguliarzer(dtf) * 0.001)
    dnet.zero_grad()
    dloss.backward()
    arrays.update({"dense.label": dlabel, "dense.cls_target": dcls, "dense.seg_target": dseg.astype(np.int8),
                   "dense.cls_logits": net1.detach().numpy(), "dense.seg_logp": net2.detach().numpy(),
                   "dense.loss": np.float64(dloss.item()),
                   "dense.keep": torch.where(seen["x"] != 0, seen["y"] != 0, torch.ones_like(seen["x"], dtype=torch.bool)).numpy().astype(np.uint8)})
    for name, p in dnet.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        arrays["dense.grad." + name] = (g if g.size <= 4096 else g[::63]).astype(np.float32)
    print(f"PointNetDenseCls: loss {dloss.item():.5f}")
    path = os.path.join(OUT, "train_pointnet_seg_seeded.npz")
    np.savez_compressed(path, **arrays)
    print(f"PointNetSeg: loss {loss.item():.5f}; wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB")


def main():
    torch.set_num_threads(8)
    ref = import_reference()
    if "--cls-only" in sys.argv:
        cls_nets(ref)
        return
    if "--pointnet-only" in sys.argv:
        pointnet_seg(ref)
        return
    blocks(ref)
    cls_nets(ref)
    pointnet_seg(ref)
    net = ref["pointnet2"].PointNet2SemSeg(CLASSES, feature_dims=1)
    sd = torch.load(os.path.join(REF, "checkpoints", CKPT), map_location="cpu")
    net.load_state_dict({k[len("module."):]: v for k, v in sd.items()})
    net.train()
    pts = syn.kitti_batch(B, N, config=5)
    target = np.random.default_rng(5000).integers(0, CLASSES, size=(B, N)).astype(np.int64)
    seen = {}
    net.drop1.register_forward_hook(lambda m, i, o: seen.update(x=i[0].detach().clone(), y=o.detach().clone()))
    torch.manual_seed(0)
    starts = [torch.randint(0, n, (B,), dtype=torch.long).numpy() for n in (N, 1024, 256, 64)]
    torch.manual_seed(0)
    logits = net(torch.from_numpy(pts))
    loss = torch.nn.CrossEntropyLoss()(logits.transpose(2, 1), torch.from_numpy(target))
    net.zero_grad()
    loss.backward()
    x, y = seen["x"], seen["y"]                       # [B,128,N] channel-major
    keep = torch.where(x != 0, (y != 0), torch.ones_like(x, dtype=torch.bool))
    keep_rows = keep.permute(0, 2, 1).reshape(B * N, -1).numpy()
    arrays = {
        "input_checksum": np.array(syn.checksum(pts)),
        "target": target.astype(np.int8),
        "starts": np.stack(starts).astype(np.int32),
        "keep_bits": np.packbits(keep_rows, axis=1),
        "loss": np.float64(loss.item()),
        "logp": logits.detach().numpy(),
    }
    for name, p in net.named_parameters():
        g = p.grad.detach().numpy().reshape(-1)
        if g.size <= FULL_MAX:
            arrays["grad." + name] = g.astype(np.float32)
        else:
            arrays["gradsub." + name] = g[::STRIDE].astype(np.float32)
        arrays["gradstat." + name] = np.array([g.astype(np.float64).sum(), np.sqrt((g.astype(np.float64) ** 2).sum())])
    for name, b in net.named_buffers():
        arrays["buffer." + name] = b.detach().numpy()
    path = os.path.join(OUT, "train_step_ckpt.npz")
    np.savez_compressed(path, **arrays)
    print(f"wrote {path}: {os.path.getsize(path) / 1e6:.2f} MB, loss {loss.item():.6f}, "
          f"kept fraction {keep_rows.mean():.4f}")


if __name__ == "__main__":
    main()
