"""CPU oracle of the TRAINING step and the evaluation metrics.

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs -- never by the product.

A numpy restatement (float64 arithmetic; the geometry -- FPS, ball query, 3-NN -- comes from the bit-exact C oracle) of
what the reference does in one iteration of its training loop, pcdseg.py:157-186:

    model.train(); logits = model(points)                 PointNet2SemSeg.forward, pointnet2.py:159-176, with
                                                          BatchNorm on batch statistics and Dropout(0.5) active
    loss = nn.CrossEntropyLoss()(logits.transpose(2,1), target)     (on the log-probabilities, pcdseg.py:177-178)
    loss.backward(); optimizer.step()                     torch.optim.Adam(lr, (0.9, 0.999), 1e-8, weight_decay=1e-4)

and of test_kitti_semseg (pcdseg.py:58-97).  Parity status: PINNED -- tests/golden/train_step_ckpt.npz holds the loss,
log-probabilities, every parameter gradient and the updated BatchNorm buffers produced by the reference itself, and
train_blocks_seeded.npz one set-abstraction and one feature-propagation block (oracle/gen_golden_train.py imports
/root/reference/model and runs its autograd on the CPU); tests/test_train_oracle.py checks this file against them, and
adam_step against torch.optim.Adam.  (The classification / part-segmentation nets and PointNetSeg have no restatement here:
their CUDA training path is checked directly against the reference's autograd fixtures, train_cls_seeded.npz and
train_pointnet_seg_seeded.npz.)

The dropout mask is an INPUT here (the reference draws it from torch's generator; the golden generator records it).
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence

import numpy as np

from . import oracle as orc

F64 = np.float64
SA_CFG = [("sa1", 1024, 0.1, 32), ("sa2", 256, 0.2, 32), ("sa3", 64, 0.4, 32), ("sa4", 16, 0.8, 32)]   # pointnet2.py:145-148
BN_EPS, BN_MOMENTUM = 1e-5, 0.1


# --------------------------------------------------------------------------- layers
def _w(sd, conv):
    w = np.asarray(sd[conv + ".weight"], dtype=F64)
    return w.reshape(w.shape[0], w.shape[1])


def conv_bn_relu_fwd(sd, conv: str, bn: Optional[str], x: np.ndarray, new_buffers: Optional[dict]):
    """relu(bn(conv(x))) with batch statistics over the rows (F.relu(bn(conv(.))), pointnet_util.py:197, :312).
    Returns (z, cache).  new_buffers collects running_mean / running_var / num_batches_tracked after the update."""
    w = _w(sd, conv)
    y = x @ w.T + np.asarray(sd[conv + ".bias"], dtype=F64)
    if bn is None:
        return y, (x, y, None)
    n = y.shape[0]
    mu = y.mean(0)
    var = y.var(0)                                  # biased: what normalisation uses
    invstd = 1.0 / np.sqrt(var + BN_EPS)
    xhat = (y - mu) * invstd
    g, b = np.asarray(sd[bn + ".weight"], dtype=F64), np.asarray(sd[bn + ".bias"], dtype=F64)
    z = np.maximum(xhat * g + b, 0.0)
    if new_buffers is not None:                     # nn.BatchNorm: momentum 0.1, unbiased variance into the running one
        new_buffers[bn + ".running_mean"] = (1 - BN_MOMENTUM) * np.asarray(sd[bn + ".running_mean"], F64) + BN_MOMENTUM * mu
        new_buffers[bn + ".running_var"] = ((1 - BN_MOMENTUM) * np.asarray(sd[bn + ".running_var"], F64)
                                            + BN_MOMENTUM * var * (n / max(1, n - 1)))
        new_buffers[bn + ".num_batches_tracked"] = np.asarray(sd[bn + ".num_batches_tracked"]) + 1
    return z, (x, y, (xhat, invstd, g, z))


def conv_bn_relu_bwd(sd, conv: str, bn: Optional[str], cache, dz: np.ndarray, grads: dict, need_dx: bool = True):
    x, y, bnc = cache
    if bnc is not None:
        xhat, invstd, g, z = bnc
        gz = dz * (z > 0)
        grads[bn + ".weight"] = (gz * xhat).sum(0)
        grads[bn + ".bias"] = gz.sum(0)
        dy = g * invstd * (gz - gz.mean(0) - xhat * (gz * xhat).mean(0))
    else:
        dy = dz
    w = _w(sd, conv)
    grads[conv + ".weight"] = (dy.T @ x).reshape(np.asarray(sd[conv + ".weight"]).shape)
    grads[conv + ".bias"] = dy.sum(0)
    return dy @ w if need_dx else None


def _nlayers(sd, prefix):
    n = 0
    while f"{prefix}.{n}.weight" in sd:
        n += 1
    return n


# --------------------------------------------------------------------------- blocks
def set_abstraction_fwd(sd, name, npoint, radius, nsample, xyz, points, start, new_buffers):
    """PointNetSetAbstraction.forward (pointnet_util.py:175-201), train mode, point-major."""
    fps_idx = orc.farthest_point_sample(xyz, npoint, start)
    new_xyz = orc.index_points(xyz, fps_idx)
    gidx = orc.query_ball_point(radius, nsample, xyz, new_xyz)
    g = orc.group(xyz, None if points is None else points.astype(np.float32), new_xyz, gidx, msg_order=False)
    B, S, K, C = g.shape
    if points is not None:                          # keep the feature channels in full precision
        g = g.astype(F64)
        g[..., 3:] = np.take_along_axis(points[:, :, None, :], gidx.reshape(B, S * K, 1, 1), axis=1).reshape(B, S, K, -1)
    h = g.reshape(B * S * K, C).astype(F64)
    caches = []
    for i in range(_nlayers(sd, f"{name}.mlp_convs")):
        h, c = conv_bn_relu_fwd(sd, f"{name}.mlp_convs.{i}", f"{name}.mlp_bns.{i}", h, new_buffers)
        caches.append(c)
    h = h.reshape(B * S, K, -1)
    am = h.argmax(1)                                # first maximum (torch.max on CPU)
    pooled = np.take_along_axis(h, am[:, None, :], axis=1)[:, 0, :]
    return new_xyz, pooled.reshape(B, S, -1), (caches, gidx, am, (B, xyz.shape[1], S, K))


def set_abstraction_bwd(sd, name, cache, dpooled, grads, need_dpoints):
    caches, gidx, am, (B, N, S, K) = cache
    Cc = dpooled.shape[-1]
    dz = np.zeros((B * S, K, Cc), dtype=F64)
    np.put_along_axis(dz, am[:, None, :], dpooled.reshape(B * S, 1, Cc), axis=1)
    dz = dz.reshape(B * S * K, Cc)
    for i in range(len(caches) - 1, -1, -1):
        dz = conv_bn_relu_bwd(sd, f"{name}.mlp_convs.{i}", f"{name}.mlp_bns.{i}", caches[i], dz, grads,
                              need_dx=(i > 0 or need_dpoints))
    if not need_dpoints:
        return None
    D = dz.shape[1] - 3
    dpts = np.zeros((B, N, D), dtype=F64)
    flat = gidx.reshape(B, S * K)
    for b in range(B):
        np.add.at(dpts[b], flat[b], dz.reshape(B, S * K, -1)[b, :, 3:])
    return dpts


def feature_propagation_fwd(sd, name, xyz1, xyz2, points1, points2, new_buffers):
    """PointNetFeaturePropagation.forward (pointnet_util.py:275-313), train mode, point-major."""
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    idx, w32, _, _ = orc.three_nn(xyz1, xyz2)
    w = w32.astype(F64)
    interp = np.zeros((B, N, points2.shape[2]), dtype=F64)
    for b in range(B):
        for k in range(3):
            interp[b] += points2[b, idx[b, :, k], :] * w[b, :, k:k + 1]
    h = interp if points1 is None else np.concatenate([points1, interp], -1)
    D1 = 0 if points1 is None else points1.shape[2]
    h = h.reshape(B * N, -1)
    caches = []
    for i in range(_nlayers(sd, f"{name}.mlp_convs")):
        h, c = conv_bn_relu_fwd(sd, f"{name}.mlp_convs.{i}", f"{name}.mlp_bns.{i}", h, new_buffers)
        caches.append(c)
    return h.reshape(B, N, -1), (caches, idx, w, (B, N, S, D1))


def feature_propagation_bwd(sd, name, cache, dout, grads):
    caches, idx, w, (B, N, S, D1) = cache
    dz = dout.reshape(B * N, -1)
    for i in range(len(caches) - 1, -1, -1):
        dz = conv_bn_relu_bwd(sd, f"{name}.mlp_convs.{i}", f"{name}.mlp_bns.{i}", caches[i], dz, grads)
    dz = dz.reshape(B, N, -1)
    dp1 = dz[:, :, :D1] if D1 else None
    di = dz[:, :, D1:]
    dp2 = np.zeros((B, S, di.shape[2]), dtype=F64)
    for b in range(B):
        for k in range(3):
            np.add.at(dp2[b], idx[b, :, k], di[b] * w[b, :, k:k + 1])
    return dp1, dp2


# --------------------------------------------------------------------------- the step
def _log_softmax(x):
    m = x.max(-1, keepdims=True)
    return x - m - np.log(np.exp(x - m).sum(-1, keepdims=True))


def semseg_train_step(sd: Dict[str, np.ndarray], points: np.ndarray, target: np.ndarray, starts: Sequence[Sequence[int]],
                      keep_mask: np.ndarray, p_drop: float = 0.5):
    """One forward + loss + backward of PointNet2SemSeg in train mode.
    points [B, 3+fd, N] channel-major; target [B, N]; starts = the four FPS start draws; keep_mask [B*N, 128] (0/1).
    Returns dict(loss, logp [B,N,k], grads {state-dict name: ndarray}, buffers {name: updated BN buffer})."""
    pts = np.asarray(points, dtype=np.float32).transpose(0, 2, 1)
    xyz, feat = np.ascontiguousarray(pts[:, :, :3]), pts[:, :, 3:].astype(F64)
    B, N, _ = xyz.shape
    buffers: dict = {}
    lx, lf, sac = [xyz], [feat], []
    for (name, npoint, radius, nsample), st in zip(SA_CFG, starts):
        nx, nf, c = set_abstraction_fwd(sd, name, npoint, radius, nsample, lx[-1], lf[-1], st, buffers)
        lx.append(nx)
        lf.append(nf)
        sac.append(c)
    f3, c4 = feature_propagation_fwd(sd, "fp4", lx[3], lx[4], lf[3], lf[4], buffers)
    f2, c3 = feature_propagation_fwd(sd, "fp3", lx[2], lx[3], lf[2], f3, buffers)
    f1, c2 = feature_propagation_fwd(sd, "fp2", lx[1], lx[2], lf[1], f2, buffers)
    f0, c1 = feature_propagation_fwd(sd, "fp1", lx[0], lx[1], None, f1, buffers)
    z1, hc = conv_bn_relu_fwd(sd, "conv1", "bn1", f0.reshape(B * N, -1), buffers)          # pointnet2.py:172
    keep = np.asarray(keep_mask, dtype=F64).reshape(B * N, -1) / (1.0 - p_drop)
    zd = z1 * keep
    logits, oc = conv_bn_relu_fwd(sd, "conv2", None, zd, None)                              # :173
    logp = _log_softmax(logits)                                                             # :174
    # nn.CrossEntropyLoss on the log-probabilities (pcdseg.py:177-178): log_softmax once more, then mean NLL
    t = np.asarray(target).reshape(-1)
    lp2 = _log_softmax(logp)
    R = B * N
    loss = -lp2[np.arange(R), t].mean()
    dlp2 = np.zeros_like(lp2)
    dlp2[np.arange(R), t] = -1.0 / R
    dlogp = dlp2 - np.exp(lp2) * dlp2.sum(-1, keepdims=True)
    dlogits = dlogp - np.exp(logp) * dlogp.sum(-1, keepdims=True)
    grads: dict = {}
    dzd = conv_bn_relu_bwd(sd, "conv2", None, oc, dlogits, grads)
    df0 = conv_bn_relu_bwd(sd, "conv1", "bn1", hc, dzd * keep, grads).reshape(B, N, -1)
    _, df1 = feature_propagation_bwd(sd, "fp1", c1, df0, grads)
    dl1, df2 = feature_propagation_bwd(sd, "fp2", c2, df1, grads)
    dl2, df3 = feature_propagation_bwd(sd, "fp3", c3, df2, grads)
    dl3, dl4 = feature_propagation_bwd(sd, "fp4", c4, df3, grads)
    dl3 = dl3 + set_abstraction_bwd(sd, "sa4", sac[3], dl4, grads, True)
    dl2 = dl2 + set_abstraction_bwd(sd, "sa3", sac[2], dl3, grads, True)
    dl1 = dl1 + set_abstraction_bwd(sd, "sa2", sac[1], dl2, grads, True)
    set_abstraction_bwd(sd, "sa1", sac[0], dl1, grads, False)
    return {"loss": float(loss), "logp": logp.reshape(B, N, -1), "grads": grads, "buffers": buffers}


def adam_step(param, grad, exp_avg, exp_avg_sq, step: int, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4):
    """torch.optim.Adam._single_tensor_adam (amsgrad off, maximize off): returns (param, exp_avg, exp_avg_sq)."""
    g = np.asarray(grad, F64) + weight_decay * np.asarray(param, F64)
    m = betas[0] * np.asarray(exp_avg, F64) + (1 - betas[0]) * g
    v = betas[1] * np.asarray(exp_avg_sq, F64) + (1 - betas[1]) * g * g
    bc1, bc2 = 1 - betas[0] ** step, 1 - betas[1] ** step
    denom = np.sqrt(v) / np.sqrt(bc2) + eps
    return np.asarray(param, F64) - (lr / bc1) * m / denom, m, v


# --------------------------------------------------------------------------- evaluation metrics
def seg_counts(logp: np.ndarray, target: np.ndarray):
    """Per-class (intersection, predicted, target) counts and correct points of one batch (pcdseg.py:72-83)."""
    k = logp.shape[-1]
    pred = logp.reshape(-1, k).argmax(-1)
    t = np.asarray(target).reshape(-1)
    inter = np.array([np.sum((pred == c) & (t == c)) for c in range(k)], dtype=np.int64)
    npred = np.bincount(pred, minlength=k).astype(np.int64)
    ntgt = np.bincount(t, minlength=k).astype(np.int64)
    return inter, npred, ntgt, int((pred == t).sum())


def test_kitti_semseg(batches, num_classes: int):
    """pcdseg.py:58-97 line by line on (logp [B,N,k], target [B,N]) pairs -> (acc, miou, categorical_iou)."""
    ious = np.zeros((num_classes,), dtype=np.float32)
    count = np.zeros((num_classes,), dtype=np.uint32)
    count[0] = 1
    accuracy = []
    for logp, target in batches:
        pred = logp.argmax(-1)
        target = np.asarray(target).reshape(pred.shape)
        for c in range(num_classes):
            I = int(np.sum((pred == c) & (target == c)))
            U = int(np.sum((pred == c) | (target == c)))
            iou = 1 if U == 0 else I / U
            ious[c] += iou
            count[c] += 1
        accuracy.append(int((pred == target).sum()) / pred.size)
    categorical = ious / count
    return float(np.mean(accuracy)), float(np.mean(categorical[1:])), categorical
