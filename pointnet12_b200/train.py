"""Training step and evaluation metrics on the native kernels.

Reference: pcdseg.py:157-186 -- `model.train()`, `logits = model(points)`, `nn.CrossEntropyLoss()(logits.transpose(2, 1),
target)`, `loss.backward()`, `optimizer.step()` with torch.optim.Adam(lr, betas (0.9, 0.999), eps 1e-8, weight_decay 1e-4)
under nn.DataParallel -- and pcdseg.py:58-97 (test_kitti_semseg).

The drop-in contract is the reference's own: in train() mode the modules of model/pointnet_util.py and PointNet2SemSeg
return tensors that carry a grad_fn, so the reference's loop (torch loss, loss.backward(), any torch optimizer) runs
unchanged; every kernel underneath -- forward with batch-statistics BatchNorm, dropout, and the whole backward -- is ours
(csrc/train.cu + linear.cu + the sampling / grouping kernels of the inference path).  One torch.autograd.Function per
block (set abstraction, feature propagation, segmentation head): autograd only connects them.

On top of that, the pieces a B200-native training loop uses instead of the torch ones:
    cross_entropy(logp, target)    the reference's loss as one kernel (forward + gradient)
    FlatAdam(params, ...)          parameters and gradients re-pointed into ONE flat buffer each: the data-parallel gradient
                                   exchange is a single all-reduce of 3.9 MB and Adam is a single kernel launch
    SegMetrics(num_classes)        argmax + per-class I/U + accuracy accumulated on the device (no .item() per class)

Layers are computed in exact fp32 on the CUDA cores here (pn_linear_f32 / pn_grad_weight_f32); moving the training
GEMMs onto the tcgen05 chains is the next step for this row.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch

from . import ops


# ------------------------------------------------------------------------------------------------
# shared MLP with batch-statistics BatchNorm: forward keeps (input, pre-norm output, statistics) per layer
def _w2d(conv) -> torch.Tensor:
    return conv.weight.detach().reshape(conv.weight.shape[0], -1)


def mlp_forward(x: torch.Tensor, layers: Sequence[Tuple], pool_K: Optional[int] = None):
    """x [rows, cin] -> (output, saved).  layers = [(conv, bn or None, relu)]; pool_K: the last layer's activation is
    max-pooled over runs of K rows (set abstraction)."""
    saved = []
    for li, (conv, bn, relu) in enumerate(layers):
        y = ops.linear(x, _w2d(conv), conv.bias.detach() if conv.bias is not None else None, relu=False)
        st, am = None, None
        if bn is not None:
            st = ops.bn_batch_stats(y, bn)
            if pool_K and li == len(layers) - 1:
                z, am = ops.bn_act_max(y, st, pool_K, relu)
            else:
                z = ops.bn_act(y, st, relu)
        else:
            if relu or (pool_K and li == len(layers) - 1):
                raise NotImplementedError("training path: a layer without BatchNorm must be a plain linear layer")
            z = y
        saved.append((x, y, st, relu, am))
        x = z
    return x, saved


def mlp_backward(saved, layers: Sequence[Tuple], dz: torch.Tensor, pool_K: Optional[int], need_dx: bool):
    """-> (dx or None, [per layer: (dW [Co,Ci], db, dgamma, dbeta)])."""
    grads = [None] * len(layers)
    for li in range(len(layers) - 1, -1, -1):
        conv, bn, relu = layers[li]
        x, y, st, _, am = saved[li]
        dgamma = dbeta = None
        if st is not None:
            dy, dgamma, dbeta = ops.bn_act_backward(y, st, dz, relu, argmax=am, K=pool_K if am is not None else 1)
        else:
            dy = dz
        w = _w2d(conv)
        dw = torch.zeros_like(w)
        db = torch.zeros((w.shape[0],), dtype=torch.float32, device=w.device) if conv.bias is not None else None
        ops.grad_weight(dy, x[:, :w.shape[1]] if x.shape[1] != w.shape[1] else x, dw, db)
        grads[li] = (dw, db, dgamma, dbeta)
        if li > 0 or need_dx:
            dz = ops.linear(dy, ops.transpose(w), None, relu=False)
        else:
            dz = None
    return dz, grads


def _layer_params(layers) -> List[torch.Tensor]:
    out = []
    for conv, bn, _ in layers:
        out.append(conv.weight)
        if conv.bias is not None:
            out.append(conv.bias)
        if bn is not None:
            out += [bn.weight, bn.bias]
    return out


def _layer_grads(layers, grads) -> List[torch.Tensor]:
    out = []
    for (conv, bn, _), (dw, db, dgamma, dbeta) in zip(layers, grads):
        out.append(dw.view_as(conv.weight))
        if conv.bias is not None:
            out.append(db)
        if bn is not None:
            out += [dgamma, dbeta]
    return out


def _rows_of(grad: torch.Tensor, C: int) -> torch.Tensor:
    """An incoming gradient [B, n, C] (any strides) as a contiguous [B*n, C] row matrix."""
    return grad.contiguous().view(-1, C)


# ------------------------------------------------------------------------------------------------
class SetAbstractionFn(torch.autograd.Function):
    """Grouping + shared MLP (batch-stat BN) + max over nsample of PointNetSetAbstraction.forward
    (model/pointnet_util.py:187-199).  Sampling / ball query carry no gradient and are done by the caller."""

    @staticmethod
    def forward(ctx, mod, xyz_pm, pts_pm, new_xyz, idx, *params):
        B, S, K = idx.shape
        D = pts_pm.shape[2] if pts_pm is not None else 0
        layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
        grouped = ops.group(xyz_pm, pts_pm, new_xyz, idx, msg_order=False, pad4=True)          # [B,S,K,ld]
        x0 = grouped.view(B * S * K, -1)[:, :3 + D]
        pooled, saved = mlp_forward(x0, layers, pool_K=K)
        ctx.mod, ctx.saved, ctx.idx, ctx.geom = mod, saved, idx, (B, xyz_pm.shape[1], S, K, D)
        return pooled.view(B, S, -1)

    @staticmethod
    def backward(ctx, dpooled):
        mod = ctx.mod
        B, N, S, K, D = ctx.geom
        layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
        need_dx = D > 0 and ctx.needs_input_grad[2]
        dx0, grads = mlp_backward(ctx.saved, layers, _rows_of(dpooled, dpooled.shape[-1]), K, need_dx)
        dpts = ops.group_backward(dx0, 3, D, ctx.idx, N) if need_dx else None
        ctx.saved = None
        return (None, None, dpts, None, None, *_layer_grads(layers, grads))


class FeaturePropagationFn(torch.autograd.Function):
    """Interpolation + skip concat + shared MLP (batch-stat BN) of PointNetFeaturePropagation.forward
    (model/pointnet_util.py:301-312); the 3-NN search carries no gradient and is done by the caller."""

    @staticmethod
    def forward(ctx, mod, p1, p2, idx, w, *params):
        B, N, _ = idx.shape
        layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
        x0 = ops.three_interpolate(p1, p2, idx, w).view(B * N, -1)
        out, saved = mlp_forward(x0, layers)
        ctx.mod, ctx.saved, ctx.nn = mod, saved, (idx, w)
        ctx.geom = (B, N, p2.shape[1], p1.shape[2] if p1 is not None else 0, p2.shape[2])
        return out.view(B, N, -1)

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        B, N, S, D1, D2 = ctx.geom
        layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
        dx0, grads = mlp_backward(ctx.saved, layers, _rows_of(dout, dout.shape[-1]), None, True)
        dp1, dp2 = ops.three_interpolate_backward(dx0, D1, D2, ctx.nn[0], ctx.nn[1], S)
        ctx.saved = None
        return (None, dp1 if ctx.needs_input_grad[1] else None, dp2 if ctx.needs_input_grad[2] else None, None, None,
                *_layer_grads(layers, grads))


class SegHeadFn(torch.autograd.Function):
    """conv1 -> bn1 -> relu -> dropout -> conv2 -> log_softmax of pointnet2.py:172-175 in training mode."""

    @staticmethod
    def forward(ctx, net, feat, mask, seed_offset, *params):
        B, N, C = feat.shape
        layers = [(net.conv1, net.bn1, True)]
        z1, saved = mlp_forward(feat.contiguous().view(B * N, C), layers)
        p = float(net.drop1.p)
        if p > 0.0 or mask is not None:
            zd, mask = ops.dropout(z1, p, seed_offset=seed_offset, mask=mask)
        else:
            zd, mask = z1, None
        logits = ops.linear(zd, _w2d(net.conv2), net.conv2.bias.detach(), relu=False)
        logp = ops.log_softmax(logits)
        ctx.net, ctx.saved, ctx.tail = net, saved, (zd, mask, logp, p)
        return logp.view(B, N, -1)

    @staticmethod
    def backward(ctx, dlogp):
        net = ctx.net
        zd, mask, logp, p = ctx.tail
        k = logp.shape[1]
        dlogits = ops.log_softmax_backward(_rows_of(dlogp, k), logp)
        w2 = _w2d(net.conv2)
        dw2, db2 = torch.zeros_like(w2), torch.zeros((k,), dtype=torch.float32, device=w2.device)
        ops.grad_weight(dlogits, zd, dw2, db2)
        dzd = ops.linear(dlogits, ops.transpose(w2), None, relu=False)
        dz1 = ops.dropout(dzd, p, mask=mask)[0] if mask is not None else dzd
        layers = [(net.conv1, net.bn1, True)]
        dfeat, grads = mlp_backward(ctx.saved, layers, dz1, None, True)
        B, N = dlogp.shape[0], dlogp.shape[1]
        ctx.saved = ctx.tail = None
        return (None, dfeat.view(B, N, -1), None, None, *_layer_grads(layers, grads), dw2.view_as(net.conv2.weight), db2)


# ------------------------------------------------------------------------------------------------
# train-mode forwards called by the modules (model/pointnet_util.py, model/pointnet2.py)
def set_abstraction_train(mod, xyz: torch.Tensor, points: Optional[torch.Tensor], start_idx=None):
    if mod.group_all:
        raise NotImplementedError("training path: group_all levels are not built yet (PointNet2SemSeg has none)")
    xyz_pm = xyz.detach().permute(0, 2, 1)
    pts_pm = points.permute(0, 2, 1) if points is not None else None
    with torch.no_grad():
        new_xyz, idx = mod.geometry(xyz_pm, start_idx)
    layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
    pooled = SetAbstractionFn.apply(mod, xyz_pm, pts_pm, new_xyz, idx, *_layer_params(layers))
    return new_xyz.permute(0, 2, 1), pooled.permute(0, 2, 1)


def feature_propagation_train(mod, xyz1, xyz2, points1, points2):
    with torch.no_grad():
        idx, w = mod.geometry(xyz1.detach().permute(0, 2, 1), xyz2.detach().permute(0, 2, 1))
    p1 = points1.permute(0, 2, 1) if points1 is not None else None
    layers = [(c, b, True) for c, b in zip(mod.mlp_convs, mod.mlp_bns)]
    out = FeaturePropagationFn.apply(mod, p1, points2.permute(0, 2, 1), idx, w, *_layer_params(layers))
    return out.permute(0, 2, 1)


_SEED = {"next": 0}


def dropout_seed(device) -> torch.Tensor:
    """{seed, offset} for the Philox stream of one dropout call: the seed comes from torch's CPU generator state
    (so torch.manual_seed makes runs repeatable), the offset counts the calls."""
    _SEED["next"] += 1
    return torch.tensor([torch.initial_seed() & 0x7FFFFFFFFFFFFFFF, _SEED["next"]], dtype=torch.int64).to(device, non_blocking=True)


def semseg_forward_train(net, points: torch.Tensor, fps_starts=None, dropout_mask: Optional[torch.Tensor] = None,
                         seed_offset: Optional[torch.Tensor] = None) -> torch.Tensor:
    """PointNet2SemSeg.forward (pointnet2.py:159-176) in train() mode -> log-probabilities [B, N, classes] with grad_fn.
    dropout_mask (extension): uint8 [B*N, 128] keep-mask to use instead of drawing one (parity tests)."""
    ops._need_cuda(points, "points")
    xyz = points[:, :3, :]
    feats = points[:, 3:, :] if points.shape[1] > 3 else None
    sa = [net.sa1, net.sa2, net.sa3, net.sa4]
    xs, fs = [xyz], [feats]
    for i, m in enumerate(sa):
        nx, nf = set_abstraction_train(m, xs[-1], fs[-1], None if fps_starts is None else fps_starts[i])
        xs.append(nx)
        fs.append(nf)
    up = feature_propagation_train(net.fp4, xs[3], xs[4], fs[3], fs[4])
    up = feature_propagation_train(net.fp3, xs[2], xs[3], fs[2], up)
    up = feature_propagation_train(net.fp2, xs[1], xs[2], fs[1], up)
    up = feature_propagation_train(net.fp1, xs[0], xs[1], None, up)
    if dropout_mask is None and seed_offset is None and net.drop1.p > 0:
        seed_offset = dropout_seed(points.device)
    params = [net.conv1.weight, net.conv1.bias, net.bn1.weight, net.bn1.bias, net.conv2.weight, net.conv2.bias]
    return SegHeadFn.apply(net, up.permute(0, 2, 1), dropout_mask, seed_offset, *params)


# ------------------------------------------------------------------------------------------------
class CrossEntropyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logp, target):
        B, N, k = logp.shape
        loss, dx = ops.cross_entropy(logp.contiguous().view(B * N, k), target)
        ctx.dx, ctx.shape = dx, (B, N, k)
        return loss

    @staticmethod
    def backward(ctx, g):
        return (ctx.dx * g).view(ctx.shape), None


def cross_entropy(logp: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
    """nn.CrossEntropyLoss()(logp.transpose(2, 1), target) of pcdseg.py:177-178 for logp [B, N, classes], target [B, N]
    (or [B, N, 1]): one kernel computes the loss and its gradient."""
    return CrossEntropyFn.apply(logp, target)


# ------------------------------------------------------------------------------------------------
class FlatAdam:
    """torch.optim.Adam semantics (pcdseg.py:136-141) over ONE flat parameter buffer.

    The module's parameters are re-pointed into `self.flat` and their .grad into `self.grad` (views), so a backward pass
    fills one contiguous gradient: the data-parallel exchange is one all-reduce and the update one kernel launch.
    `param_groups[0]['lr']` may be changed between steps like the reference's decay does (pcdseg.py:160-164)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatAdam: no parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty((n,), dtype=torch.float32, device=dev)
        self.grad = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        o = 0
        for p in self.params:
            k = p.numel()
            self.flat[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)
            o += k
        self.param_groups = [{"lr": lr, "betas": betas, "eps": eps, "weight_decay": weight_decay, "params": self.params}]
        self.steps = 0

    def zero_grad(self, set_to_none: bool = False):
        self.grad.zero_()
        o = 0
        for p in self.params:              # keep .grad pointing into the flat buffer (autograd accumulates in place)
            k = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.grad.data_ptr() + 4 * o:
                p.grad = self.grad[o:o + k].view(p.shape)
            o += k

    def all_reduce(self, group=None):
        """Sum the flat gradient over the data-parallel ranks (one NCCL all-reduce of numel*4 bytes); the division by
        the world size is folded into the update (grad_scale)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
            return 1.0 / dist.get_world_size(group)
        return 1.0

    def step(self, grad_scale: float = 1.0):
        g = self.param_groups[0]
        self.steps += 1
        ops.adam_step(self.flat, self.grad, self.exp_avg, self.exp_avg_sq, self.steps, g["lr"], g["betas"], g["eps"],
                      g["weight_decay"], grad_scale)
        self.mark_updated()

    def mark_updated(self):
        """The kernel wrote the parameters through the flat buffer: bump their version counters so that caches keyed on
        them (the eval path's folded / packed weights) notice."""
        torch.autograd.graph.increment_version(self.params)


# ------------------------------------------------------------------------------------------------
class SegMetrics:
    """test_kitti_semseg (pcdseg.py:58-97) with the per-batch, per-class bookkeeping on the device.

        m = SegMetrics(num_classes, device)
        for points, target in loader:  m.update(model(points), target)
        acc, miou, per_class = m.result()          # the only host synchronisation

    Bit-compatible with the reference's arithmetic: iou = 1 if U == 0 else I/U per batch and class (double division,
    added into an fp32 array), count starts at 1 for class 0 (pcdseg.py:61), miou = mean(ious[1:] / count[1:])."""

    def __init__(self, num_classes: int, device):
        self.k = int(num_classes)
        self.ious = torch.zeros((self.k,), dtype=torch.float32, device=device)
        self.count = torch.zeros((self.k,), dtype=torch.int32, device=device)
        self.count[0] = 1
        self.acc_sum = torch.zeros((1,), dtype=torch.float64, device=device)
        self.batches = torch.zeros((1,), dtype=torch.int64, device=device)

    def update(self, logp: torch.Tensor, target: torch.Tensor) -> None:
        from . import _native as nv

        counts = ops.seg_metrics(logp, target)
        points = target.numel()
        with ops._on_device(counts):
            nv.call("pn_seg_metrics_accumulate", counts.data_ptr(), self.k, points, self.ious.data_ptr(), self.count.data_ptr(),
                    self.acc_sum.data_ptr(), self.batches.data_ptr(), ops._stream())

    def result(self):
        import numpy as np

        ious = self.ious.cpu().numpy()
        count = self.count.cpu().numpy().astype(np.uint32)
        categorical = ious / count
        acc = float(self.acc_sum.item() / max(1, int(self.batches.item())))
        return acc, float(np.mean(categorical[1:])), categorical
