"""GPU parity of the training step and the evaluation metrics (f-2), through the C ABI.

Three layers of evidence: every kernel of csrc/train.cu against a numpy restatement; one set-abstraction and one
feature-propagation block (forward, input gradients, parameter gradients, BatchNorm buffers) against the reference's own
autograd (tests/golden/train_blocks_seeded.npz) and the float64 oracle, to fp32 round-off; and the whole PointNet2SemSeg
training iteration on the shipped checkpoint against the reference's run (train_step_ckpt.npz) inside the band the
reference's own float32-vs-float64 difference defines (tests/test_train_oracle.py explains the band).
"""
import numpy as np
import pytest
import torch

from oracle import train_oracle as tor
from test_train_oracle import block_inputs, check_step_against_golden, rl2, seeded_block, step_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda", 0)


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


@pytest.fixture(params=["bf16x3", "fp32"])
def gemm_mode(request, dev):
    """Both GEMM engines of the training step: tensor cores (3-pass split bf16, fp32 parity) and CUDA cores (exact fp32)."""
    from pointnet12_b200 import ops

    old = ops.set_mlp_mode(request.param)
    yield request.param
    ops.set_mlp_mode(old)


# ------------------------------------------------------------------------------------------------ kernels
@pytest.mark.parametrize("rows,C,ld", [(4096, 32, 32), (1000, 67, 68), (70001, 128, 128), (33, 19, 19), (512, 512, 512)])
def test_bn_forward_kernels(dev, rows, C, ld):
    from pointnet12_b200 import ops

    rng = np.random.default_rng(rows + C)
    y = (rng.standard_normal((rows, C)) * rng.uniform(0.1, 3, C) + rng.uniform(-20, 20, C)).astype(np.float32)
    buf = torch.zeros((rows, ld), device=dev)
    buf[:, :C] = T(y, dev)
    yv = buf[:, :C]
    bn = torch.nn.BatchNorm1d(C).to(dev)
    with torch.no_grad():
        bn.weight.copy_(T(rng.uniform(0.5, 1.5, C).astype(np.float32), dev))
        bn.bias.copy_(T(rng.standard_normal(C).astype(np.float32), dev))
        bn.running_mean.copy_(T(rng.standard_normal(C).astype(np.float32), dev))
    rm0, rv0 = bn.running_mean.cpu().numpy().copy(), bn.running_var.cpu().numpy().copy()
    st = ops.bn_batch_stats(yv, bn)
    y64 = y.astype(np.float64)
    mu, var = y64.mean(0), y64.var(0)
    assert np.abs(st.mean.cpu().numpy() - mu).max() < 1e-5 * max(1, np.abs(mu).max())
    assert np.abs(st.invstd.cpu().numpy() * np.sqrt(var + 1e-5) - 1).max() < 1e-5
    assert np.abs(bn.running_mean.cpu().numpy() - (0.9 * rm0 + 0.1 * mu)).max() < 1e-5
    assert np.abs(bn.running_var.cpu().numpy() - (0.9 * rv0 + 0.1 * var * rows / (rows - 1))).max() < 1e-5 * max(1, var.max())
    assert int(bn.num_batches_tracked) == 1
    g, b = bn.weight.detach().cpu().numpy().astype(np.float64), bn.bias.detach().cpu().numpy().astype(np.float64)
    want = np.maximum((y64 - mu) / np.sqrt(var + 1e-5) * g + b, 0)
    z = ops.bn_act(yv, st, relu=True).cpu().numpy()
    assert np.abs(z - want).max() < 2e-5 * max(1, np.abs(want).max())
    if rows % 32 == 0:
        pooled, am = ops.bn_act_max(yv, st, 32, relu=True)
        w3 = want.reshape(rows // 32, 32, C)
        assert np.abs(pooled.cpu().numpy() - w3.max(1)).max() < 2e-5 * max(1, np.abs(want).max())
        got_am = am.cpu().numpy().astype(np.int64)
        z3 = z.reshape(rows // 32, 32, C)
        assert np.array_equal(got_am, z3.argmax(1))                      # first maximum of what the kernel itself computes
        # backward through the pooling: gradient routed to the arg-max row only
        dpool = rng.standard_normal((rows // 32, C)).astype(np.float32)
        dy, dg, db = ops.bn_act_backward(yv, st, T(dpool, dev), relu=True, argmax=am, K=32)
        dz = np.zeros((rows // 32, 32, C))
        np.put_along_axis(dz, got_am[:, None, :], dpool[:, None, :].astype(np.float64), axis=1)
        _check_bn_backward(y64, mu, var, g, z > 0, dz.reshape(rows, C), dy, dg, db)
    dzr = rng.standard_normal((rows, C)).astype(np.float32)
    dy, dg, db = ops.bn_act_backward(yv, st, T(dzr, dev), relu=True)
    _check_bn_backward(y64, mu, var, g, z > 0, dzr.astype(np.float64), dy, dg, db)


def _check_bn_backward(y64, mu, var, g, mask, dz, dy, dg, db):
    invstd = 1 / np.sqrt(var + 1e-5)
    xhat = (y64 - mu) * invstd
    gz = dz * mask
    want = g * invstd * (gz - gz.mean(0) - xhat * (gz * xhat).mean(0))
    assert rl2(dy.cpu().numpy(), want) < 2e-5
    assert rl2(dg.cpu().numpy(), (gz * xhat).sum(0)) < 2e-5
    assert rl2(db.cpu().numpy(), gz.sum(0)) < 2e-5


@pytest.mark.parametrize("rows,cout,cin,ldx", [(8192, 64, 67, 68), (100003, 128, 128, 128), (777, 19, 128, 128), (4096, 32, 4, 4),
                                               (2048, 256, 320, 320), (50, 512, 259, 260)])
@pytest.mark.parametrize("engine", ["fp32", "tc"])
def test_grad_weight_and_input_gradient(dev, rows, cout, cin, ldx, engine):
    from pointnet12_b200 import ops

    rng = np.random.default_rng(rows)
    dy = rng.standard_normal((rows, cout)).astype(np.float32)
    x = rng.standard_normal((rows, cin)).astype(np.float32)
    w = rng.standard_normal((cout, cin)).astype(np.float32)
    xb = torch.zeros((rows, ldx), device=dev)
    xb[:, :cin] = T(x, dev)
    dw = torch.zeros((cout, cin), device=dev)
    db = torch.zeros((cout,), device=dev)
    ops.grad_weight(T(dy, dev), xb[:, :cin], dw, db, engine=engine)
    assert rl2(dw.cpu().numpy(), dy.astype(np.float64).T @ x.astype(np.float64)) < 2e-5
    assert rl2(db.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5
    ops.grad_weight(T(dy, dev), xb[:, :cin], dw, db, engine=engine)       # accumulates
    assert rl2(dw.cpu().numpy(), 2 * (dy.astype(np.float64).T @ x.astype(np.float64))) < 2e-5
    wt = ops.transpose(T(w, dev))
    assert np.array_equal(wt.cpu().numpy(), w.T)
    want_dx = dy.astype(np.float64) @ w.astype(np.float64)
    dx = ops.linear(T(dy, dev), wt, None, relu=False)
    assert rl2(dx.cpu().numpy(), want_dx) < 1e-5
    if engine == "tc":        # the input-gradient GEMM on the tensor cores: W packed transposed, no materialised W^T
        from pointnet12_b200.train import _gemm

        old = ops.set_mlp_mode("bf16x3")
        try:
            dx_tc = _gemm(T(dy, dev), T(w, dev), None, transposed=True)
            y_tc = _gemm(xb[:, :cin], T(w, dev), None)
        finally:
            ops.set_mlp_mode(old)
        assert rl2(dx_tc.cpu().numpy(), want_dx) < 2e-5, rl2(dx_tc.cpu().numpy(), want_dx)
        assert rl2(y_tc.cpu().numpy(), x.astype(np.float64) @ w.astype(np.float64).T) < 2e-5


@pytest.mark.parametrize("rows,cin,cout,ldx,transposed", [
    (1000, 67, 64, 68, False), (70001, 128, 128, 128, False), (4096, 4, 32, 4, False), (333, 131, 128, 132, False),
    (5000, 128, 256, 128, False), (2048, 128, 19, 128, False), (9000, 64, 67, 64, True), (130, 256, 128, 256, True),
    (127, 32, 32, 32, False)])
def test_fused_train_gemm(dev, rows, cin, cout, ldx, transposed):
    """pn_train_gemm_bf16x3: y = relu(x*scale+shift) @ W^T + b with the batch statistics of y from the epilogue, against
    float64 numpy; then the weight-gradient kernel applying the same transform to its x operand."""
    from pointnet12_b200 import ops

    assert ops.train_gemm_supported(cin, cout)
    rng = np.random.default_rng(rows + cin)
    x = (rng.standard_normal((rows, cin)) * 2 + 0.5).astype(np.float32)
    w = rng.standard_normal((cout, cin)).astype(np.float32)
    b = rng.standard_normal(cout).astype(np.float32)
    xb = torch.zeros((rows, ldx), device=dev)
    xb[:, :cin] = T(x, dev)
    wt = T(w.T.copy() if transposed else w, dev)
    # plain
    y = ops.train_gemm(xb[:, :cin], wt, T(b, dev), transposed=transposed)
    want = x.astype(np.float64) @ w.astype(np.float64).T + b
    assert rl2(y.cpu().numpy(), want) < 2e-5
    # fused: input transform + output statistics
    st = ops.BatchStats()
    st.scale, st.shift = T(rng.uniform(0.5, 1.5, cin).astype(np.float32), dev), T(rng.standard_normal(cin).astype(np.float32), dev)
    st.mean = st.invstd = st.scale
    acc = torch.zeros((2, cout), dtype=torch.float64, device=dev)
    y = ops.train_gemm(xb[:, :cin], wt, T(b, dev), in_stats=st, stats_acc=acc, transposed=transposed)
    z = np.maximum(x.astype(np.float64) * st.scale.cpu().numpy().astype(np.float64) + st.shift.cpu().numpy().astype(np.float64), 0)
    want = z @ w.astype(np.float64).T + b
    assert rl2(y.cpu().numpy(), want) < 2e-5
    got = acc.cpu().numpy()
    y64 = y.cpu().numpy().astype(np.float64)
    # sums over 8 rows are formed in fp32 (the first three butterfly steps), everything beyond in fp64
    assert np.abs(got[0] - y64.sum(0)).max() <= 2e-7 * np.abs(y64).sum(0).max() + 1e-6
    assert np.abs(got[1] - (y64 ** 2).sum(0)).max() <= 2e-7 * (y64 ** 2).sum(0).max() + 1e-6
    # weight gradient with the same transform applied to its x operand
    dy = rng.standard_normal((rows, 48)).astype(np.float32)
    dw = torch.zeros((48, cin), device=dev)
    db = torch.zeros((48,), device=dev)
    ops.grad_weight(T(dy, dev), xb[:, :cin], dw, db, x_stats=st)
    assert rl2(dw.cpu().numpy(), dy.astype(np.float64).T @ z) < 2e-5
    assert rl2(db.cpu().numpy(), dy.astype(np.float64).sum(0)) < 1e-5


@pytest.mark.parametrize("rows,co,ci", [(4096, 64, 64), (70001, 128, 128), (1000, 128, 67), (333, 32, 4)])
def test_input_gradient_gemm_with_fused_bn_backward_statistics(dev, rows, co, ci):
    """pn_train_gemm_bnbwd_bf16x3: dz = dy W and, in the epilogue, sum g / sum g*xhat of the layer below."""
    from pointnet12_b200 import ops

    rng = np.random.default_rng(rows + ci)
    dy = rng.standard_normal((rows, co)).astype(np.float32)
    w = rng.standard_normal((co, ci)).astype(np.float32)
    yp = rng.standard_normal((rows, ci)).astype(np.float32)
    st = ops.BatchStats()
    st.scale, st.shift = T(rng.uniform(0.5, 1.5, ci).astype(np.float32), dev), T(rng.standard_normal(ci).astype(np.float32) * 0.3, dev)
    st.mean, st.invstd = T(rng.standard_normal(ci).astype(np.float32) * 0.1, dev), T(rng.uniform(0.5, 2, ci).astype(np.float32), dev)
    acc = torch.zeros((2, ci), dtype=torch.float64, device=dev)
    dz = ops.train_gemm_bnbwd(T(dy, dev), T(w, dev), T(yp, dev), st, acc)
    want = dy.astype(np.float64) @ w.astype(np.float64)
    assert rl2(dz.cpu().numpy(), want) < 2e-5
    sc, sh, mu, inv = (t.cpu().numpy().astype(np.float64) for t in (st.scale, st.shift, st.mean, st.invstd))
    dzg = dz.cpu().numpy().astype(np.float64)
    mask = (yp.astype(np.float32) * st.scale.cpu().numpy() + st.shift.cpu().numpy()) > 0          # fp32 like the kernel (fma vs mul+add: see below)
    g = dzg * mask
    xh = (yp.astype(np.float64) - mu) * inv
    got = acc.cpu().numpy()
    # a few elements sit within one rounding of the ReLU threshold (the kernel uses a fused multiply-add): allow their weight
    slack = 1e-5 * np.abs(dzg).sum(0).max()
    assert np.abs(got[0] - g.sum(0)).max() < slack and np.abs(got[1] - (g * xh).sum(0)).max() < 3 * slack


def test_group_and_interpolate_backward(dev):
    from pointnet12_b200 import ops

    rng = np.random.default_rng(5)
    B, N, S, K, D = 2, 500, 64, 32, 13
    idx = rng.integers(0, N, (B, S, K))
    dg = rng.standard_normal((B * S * K, 3 + D)).astype(np.float32)
    got = ops.group_backward(T(dg, dev), 3, D, T(idx, dev), N).cpu().numpy()
    want = np.zeros((B, N, D))
    for b in range(B):
        np.add.at(want[b], idx[b].reshape(-1), dg.reshape(B, S * K, -1)[b, :, 3:].astype(np.float64))
    assert rl2(got, want) < 1e-5
    D1, D2, Sc = 7, 24, 40
    nn_idx = rng.integers(0, Sc, (B, N, 3))
    wgt = rng.dirichlet(np.ones(3), (B, N)).astype(np.float32)
    dx = rng.standard_normal((B * N, D1 + D2)).astype(np.float32)
    dp1, dp2 = ops.three_interpolate_backward(T(dx, dev), D1, D2, T(nn_idx, dev), T(wgt, dev), Sc)
    assert np.array_equal(dp1.cpu().numpy(), dx.reshape(B, N, -1)[:, :, :D1])
    want2 = np.zeros((B, Sc, D2))
    for b in range(B):
        for k in range(3):
            np.add.at(want2[b], nn_idx[b, :, k], dx.reshape(B, N, -1)[b, :, D1:].astype(np.float64) * wgt[b, :, k:k + 1])
    assert rl2(dp2.cpu().numpy(), want2) < 1e-5
    _, dp2_only = ops.three_interpolate_backward(T(dx[:, D1:].copy(), dev), 0, D2, T(nn_idx, dev), T(wgt, dev), Sc)
    assert rl2(dp2_only.cpu().numpy(), want2) < 1e-5


def test_dropout(dev):
    from pointnet12_b200 import ops

    rows, C = 5000, 128
    x = torch.randn((rows, C), device=dev)
    seed = torch.tensor([1234, 1], dtype=torch.int64, device=dev)
    y, mask = ops.dropout(x, 0.5, seed_offset=seed)
    y2, mask2 = ops.dropout(x, 0.5, seed_offset=seed)
    assert torch.equal(mask, mask2) and torch.equal(y, y2)                 # a stream is a function of (seed, offset)
    _, mask3 = ops.dropout(x, 0.5, seed_offset=torch.tensor([1234, 2], dtype=torch.int64, device=dev))
    assert not torch.equal(mask, mask3)
    keep = mask.float().mean().item()
    assert abs(keep - 0.5) < 4 * 0.5 / np.sqrt(rows * C)
    assert abs(mask.float().mean(0).cpu().numpy() - 0.5).max() < 0.05 and abs(mask.float().mean(1).cpu().numpy() - 0.5).max() < 0.25
    assert torch.equal(y, torch.where(mask.bool(), x * 2.0, torch.zeros_like(x)))
    _, m25 = ops.dropout(x, 0.25, seed_offset=seed)
    assert abs(m25.float().mean().item() - 0.75) < 0.005
    given = (torch.rand((rows, C), device=dev) < 0.3).to(torch.uint8)
    y4, m4 = ops.dropout(x, 0.5, mask=given)
    assert m4 is given and torch.equal(y4, torch.where(given.bool(), x * 2.0, torch.zeros_like(x)))


def test_cross_entropy_and_log_softmax_backward(dev):
    from pointnet12_b200 import ops

    rng = np.random.default_rng(9)
    rows, C = 4099, 19
    logits = rng.standard_normal((rows, C)) * 3
    logp = (logits - np.log(np.exp(logits).sum(-1, keepdims=True))).astype(np.float32)
    tgt = rng.integers(0, C, rows)
    loss, dx = ops.cross_entropy(T(logp, dev), T(tgt, dev))
    xt = torch.from_numpy(logp.astype(np.float64)).requires_grad_(True)
    ref = torch.nn.CrossEntropyLoss()(xt, torch.from_numpy(tgt))
    ref.backward()
    assert abs(loss.item() - ref.item()) < 1e-5
    assert rl2(dx.cpu().numpy(), xt.grad.numpy()) < 1e-5
    dy = rng.standard_normal((rows, C)).astype(np.float32)
    lt = torch.from_numpy(logits).requires_grad_(True)
    torch.log_softmax(lt, -1).backward(torch.from_numpy(dy.astype(np.float64)))
    got = ops.log_softmax_backward(T(dy, dev), T(logp, dev)).cpu().numpy()
    assert rl2(got, lt.grad.numpy()) < 1e-5


def test_adam_kernel_vs_torch(dev):
    from pointnet12_b200 import ops

    rng = np.random.default_rng(3)
    n = 100003
    p0 = rng.standard_normal(n).astype(np.float32)
    ref = torch.nn.Parameter(torch.from_numpy(p0.copy()))
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    p, m, v = T(p0, dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    for step in range(1, 5):
        g = (rng.standard_normal(n) * 10.0 ** rng.integers(-6, 1, n)).astype(np.float32)
        ref.grad = torch.from_numpy(g.copy())
        opt.step()
        ops.adam_step(p, T(g, dev), m, v, step, 1e-3, (0.9, 0.999), 1e-8, 1e-4)
        assert np.abs(p.cpu().numpy() - ref.detach().numpy()).max() < 1e-6
    ops.adam_step(p, T(g * 4, dev), m, v, 5, 1e-3, (0.9, 0.999), 1e-8, 1e-4, grad_scale=0.25)   # all-reduce sum / world
    ref.grad = torch.from_numpy(g.copy())
    opt.step()
    assert np.abs(p.cpu().numpy() - ref.detach().numpy()).max() < 1e-6


def test_seg_metrics_bit_exact(dev):
    """pcdseg.py:58-97: per-class I/U per batch, fp32 accumulation of double ratios, accuracy list."""
    from pointnet12_b200.train import SegMetrics

    rng = np.random.default_rng(11)
    k = 19
    batches = []
    m = SegMetrics(k, dev)
    for i in range(3):
        logp = np.log(rng.dirichlet(np.ones(k), size=(4, 6000))).astype(np.float32)
        tgt = rng.integers(0, k if i else 7, size=(4, 6000))
        if i == 0:
            logp[..., 15:] = -50.0
        batches.append((logp, tgt))
        m.update(T(logp, dev), T(tgt, dev))
    acc, miou, cat = tor.test_kitti_semseg(batches, k)
    got_acc, got_miou, got_cat = m.result()
    assert np.array_equal(got_cat, cat)
    assert got_miou == miou
    assert abs(got_acc - acc) < 1e-15
    from pointnet12_b200 import ops

    counts, pred = ops.seg_metrics(T(batches[1][0], dev), T(batches[1][1], dev), want_pred=True)
    I, P, Tt, correct = tor.seg_counts(*batches[1])
    c = counts.cpu().numpy()
    assert np.array_equal(c[:k], I) and np.array_equal(c[k:2 * k], P) and np.array_equal(c[2 * k:3 * k], Tt) and c[3 * k] == correct
    assert np.array_equal(pred.cpu().numpy(), batches[1][0].argmax(-1))


# ------------------------------------------------------------------------------------------------ blocks
def _close_but_for_flips(got, want, strict):
    """Input gradients of a max-pooled block.  The gradient of a (group, channel) goes to the arg-max row; when the two
    largest activations of a group differ by less than the GEMM engines' rounding difference (~1e-5 relative between the
    split-bf16 tensor-core products and fp32 FMAs), the engines route it to DIFFERENT points -- an O(1) change of a few
    entries that the reference's own float32-vs-float64 runs show as well.  So: exact-fp32 engine = strict L2 bound;
    tensor-core engine = tiny median error, fewer than 1 % of the entries off by more than 1e-4 of the largest, and a loose
    L2 bound on everything."""
    if strict:
        assert rl2(got, want) < 1e-4
        return
    err = np.abs(got.astype(np.float64) - want).reshape(-1)
    scale = np.abs(want).max()
    assert np.median(err) < 1e-5 * scale and (err > 1e-4 * scale).mean() < 0.01, (np.median(err), (err > 1e-4 * scale).mean())
    assert rl2(got, want) < 2e-2


def _grads(module):
    return {n: p.grad.detach().cpu().numpy() for n, p in module.named_parameters()}


def test_sa_block_train_vs_reference(dev, golden, gemm_mode):
    from pointnet12_b200.model import pointnet_util as ours

    g = golden("train_blocks_seeded")
    xyz, f1, _, _, g_sa, _ = block_inputs()
    sa = seeded_block(lambda: ours.PointNetSetAbstraction(256, 0.2, 32, 64 + 3, [64, 64, 128], False), 4321).to(dev)
    pts = T(f1, dev).requires_grad_(True)
    new_xyz, out = sa(T(xyz, dev), pts, start_idx=T(g["sa.start"].astype(np.int64), dev))
    assert out.shape == (2, 128, 256) and out.requires_grad
    out.backward(T(g_sa, dev))
    assert rl2(out.detach().cpu().numpy(), g["sa.out"]) < 2e-5
    _close_but_for_flips(pts.grad.cpu().numpy(), g["sa.dpoints"], strict=gemm_mode == "fp32")
    # parameter gradients: sums over 16 k rows of random-sign terms, so the ~40 arg-max flips the tensor-core engine's
    # 1e-5 rounding difference causes among 65 k (group, channel) pairs show up at the 1e-3 level (see _close_but_for_flips)
    tol = 1e-4 if gemm_mode == "fp32" else 1e-2
    for n, gr in _grads(sa).items():
        ref = g["sa.grad." + n]
        if "convs" in n and n.endswith("bias"):
            assert np.abs(gr).max() < 1e-3, n                             # true gradient 0 (BatchNorm follows)
        else:
            assert rl2(gr, ref) < tol, (n, rl2(gr, ref))
    for n, b in sa.named_buffers():
        ref = g["sa.buffer." + n]
        assert np.abs(b.cpu().numpy().astype(np.float64) - ref).max() < 1e-5 * max(1.0, np.abs(ref).max()), n


def test_fp_block_train_vs_reference(dev, golden, gemm_mode):
    from pointnet12_b200.model import pointnet_util as ours

    g = golden("train_blocks_seeded")
    xyz, f1, xyz2, f2, _, g_fp = block_inputs()
    fp = seeded_block(lambda: ours.PointNetFeaturePropagation(320, [256, 128]), 4322).to(dev)
    p1, p2 = T(f1, dev).requires_grad_(True), T(f2, dev).requires_grad_(True)
    out = fp(T(xyz, dev), T(xyz2, dev), p1, p2)
    out.backward(T(g_fp, dev))
    assert rl2(out.detach().cpu().numpy(), g["fp.out"]) < 2e-5
    assert rl2(p1.grad.cpu().numpy(), g["fp.dpoints1"]) < 1e-4
    assert rl2(p2.grad.cpu().numpy(), g["fp.dpoints2"]) < 1e-4
    for n, gr in _grads(fp).items():
        if "convs" in n and n.endswith("bias"):
            assert np.abs(gr).max() < 1e-3, n
        else:
            assert rl2(gr, g["fp.grad." + n]) < 1e-4, (n, rl2(gr, g["fp.grad." + n]))


# ------------------------------------------------------------------------------------------------ the step
def _train_net(dev, ckpt_path):
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg

    net = PointNet2SemSeg(19, feature_dims=1)
    sd = torch.load(ckpt_path, map_location="cpu")
    net.load_state_dict({k[len("module."):]: v for k, v in sd.items()}, strict=True)
    return net.to(dev).train()


def test_train_step_vs_reference_and_oracle(dev, golden, ckpt_path, ckpt_state, gemm_mode):
    """The reference's own iteration, written as in pcdseg.py:166-186, on our modules."""
    from pointnet12_b200.train import cross_entropy, semseg_forward_train

    g, pts, target, starts, keep = step_inputs(golden)
    net = _train_net(dev, ckpt_path)
    mask = T(keep.astype(np.uint8), dev)
    logp = semseg_forward_train(net, T(pts, dev), fps_starts=[T(s, dev) for s in starts], dropout_mask=mask)
    assert logp.shape == (2, 2048, 19) and logp.grad_fn is not None
    loss = torch.nn.CrossEntropyLoss()(logp.transpose(2, 1), T(target, dev))       # torch's loss on our output: drop-in
    net.zero_grad()
    loss.backward()
    grads = _grads(net)
    buffers = {n: b.cpu().numpy() for n, b in net.named_buffers()}
    # band 3: the CUDA path is a third float32 evaluation order of a computation whose float32 results already differ
    # from each other by this much (see tests/test_train_oracle.py)
    check_step_against_golden(g, loss.item(), logp.detach().cpu().numpy(), grads, buffers, band=3.0)
    out = tor.semseg_train_step(ckpt_state, pts, target, starts, keep)
    assert rl2(logp.detach().cpu().numpy(), out["logp"]) < 5e-4
    assert rl2(grads["conv2.weight"], out["grads"]["conv2.weight"]) < 2e-4
    assert rl2(grads["sa1.mlp_convs.0.weight"], out["grads"]["sa1.mlp_convs.0.weight"]) < 2e-2
    # our fused loss kernel = torch's loss and gradient
    net2 = _train_net(dev, ckpt_path)
    logp2 = semseg_forward_train(net2, T(pts, dev), fps_starts=[T(s, dev) for s in starts], dropout_mask=mask)
    loss2 = cross_entropy(logp2, T(target, dev))
    loss2.backward()
    assert abs(loss2.item() - loss.item()) < 1e-5
    g2 = _grads(net2)
    assert rl2(g2["conv2.weight"], grads["conv2.weight"]) < 1e-5 and rl2(g2["fp1.mlp_convs.0.weight"], grads["fp1.mlp_convs.0.weight"]) < 1e-3


def test_reference_training_loop_runs_and_learns(dev):
    """pcdseg.py:112-186 in miniature: model.train(), torch Adam with the reference's hyper-parameters, a few iterations on
    one batch -- the loss must fall; then FlatAdam (one flat buffer, one kernel) must track torch.optim.Adam."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.train import FlatAdam, cross_entropy

    pts = T(syn.kitti_batch(4, 2048, config=5), dev)
    target = (pts[:, 2, :] > pts[:, 2, :].median()).long() + 2 * (pts[:, 3, :] > 0).long()     # a learnable labelling
    torch.manual_seed(1)
    a = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    b = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    b.load_state_dict(a.state_dict())
    a.drop1.p = b.drop1.p = 0.0                      # identical arithmetic in both replicas
    opt_a = torch.optim.Adam(a.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    opt_b = FlatAdam(b.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    losses = []
    for it in range(6):
        starts = [torch.randint(0, n, (4,), dtype=torch.long).to(dev) for n in (2048, 1024, 256, 64)]
        la = torch.nn.CrossEntropyLoss()(a(pts, fps_starts=starts).transpose(2, 1), target)
        opt_a.zero_grad()
        la.backward()
        opt_a.step()
        lb = cross_entropy(b(pts, fps_starts=starts), target)
        opt_b.zero_grad()
        lb.backward()
        opt_b.step()
        losses.append((la.item(), lb.item()))
    assert losses[-1][0] < 0.7 * losses[0][0], losses
    assert abs(losses[0][0] - losses[0][1]) < 1e-4, losses
    assert abs(losses[-1][0] - losses[-1][1]) < 0.05 * losses[-1][0], losses     # atomics order + Adam's sign-like first steps
    # eval() after training: the fused inference path must refold BatchNorm from the UPDATED parameters and running
    # statistics (both were written by our kernels through raw pointers) -- checked against the CPU oracle on b's state
    from oracle import oracle as orc

    with torch.no_grad():
        out = b.eval()(pts, fps_starts=starts)
    want = orc.pointnet2_semseg(orc.numpy_state_dict(b.state_dict()), pts.cpu().numpy(), [s.cpu().numpy() for s in starts])
    assert float(np.abs(out.cpu().numpy() - want).max() / max(1.0, np.abs(want).max())) < 1e-3


def test_graphed_train_step_matches_eager(dev):
    """GraphedTrainStep (one CUDA-graph replay per iteration) == the eager autograd iteration: same draws, same kernels."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.train import FlatAdam, GraphedTrainStep, cross_entropy

    pts = T(syn.kitti_batch(2, 4096, config=5), dev)
    target = T(np.random.default_rng(1).integers(0, 19, (2, 4096)), dev)
    torch.manual_seed(3)
    a = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    b = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    b.load_state_dict(a.state_dict())
    a.drop1.p = b.drop1.p = 0.0
    opt_a = FlatAdam(a.parameters(), lr=1e-3, weight_decay=1e-4)
    opt_b = FlatAdam(b.parameters(), lr=1e-3, weight_decay=1e-4)
    runner = GraphedTrainStep(b, opt_b)
    la, lb = [], []
    torch.manual_seed(10)
    for _ in range(4):
        loss = cross_entropy(a(pts), target)
        opt_a.zero_grad()
        loss.backward()
        opt_a.step()
        la.append(loss.item())
    torch.manual_seed(10)
    for _ in range(4):
        lb.append(runner(pts, target).item())
    # the same again with the next batch's geometry prefetched during every replay (two batches alternating)
    c = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    c.load_state_dict(a.state_dict())
    a2 = PointNet2SemSeg(19, feature_dims=1).to(dev).train()
    a2.load_state_dict(a.state_dict())
    a2.drop1.p = c.drop1.p = 0.0
    opt_a2, opt_c = FlatAdam(a2.parameters(), lr=1e-3, weight_decay=1e-4), FlatAdam(c.parameters(), lr=1e-3, weight_decay=1e-4)
    pts2 = T(syn.kitti_batch(2, 4096, config=5, first=2), dev)
    batches = [pts, pts2, pts, pts2, pts]
    le, lp = [], []
    torch.manual_seed(11)
    for i in range(4):
        loss = cross_entropy(a2(batches[i]), target)
        opt_a2.zero_grad()
        loss.backward()
        opt_a2.step()
        le.append(loss.item())
    torch.manual_seed(11)
    pre = GraphedTrainStep(c, opt_c)
    for i in range(4):
        lp.append(pre(batches[i], target, next_points=batches[i + 1]).item())
    assert abs(le[0] - lp[0]) < 1e-5 and max(abs(x - y) / x for x, y in zip(le, lp)) < 2e-2, (le, lp)
    assert abs(la[0] - lb[0]) < 1e-5, (la, lb)                # identical weights, identical draws
    assert max(abs(x - y) / x for x, y in zip(la, lb)) < 2e-2, (la, lb)     # later: fp32 atomics order through Adam
    assert opt_a.steps == opt_b.steps == 4
    assert int(b.bn1.num_batches_tracked) == 4                # the warm-up iterations of the capture were rolled back
    assert rl2(b.bn1.running_mean.cpu().numpy(), a.bn1.running_mean.cpu().numpy()) < 5e-2
    # dropout inside the graph: a fresh Philox offset per replay
    b.drop1.p = 0.5
    runner2 = GraphedTrainStep(b, opt_b)
    l1, l2 = runner2(pts, target).item(), runner2(pts, target).item()
    assert np.isfinite(l1) and np.isfinite(l2) and l1 != l2


@pytest.mark.parametrize("B,N", [(3, 1500), (1, 1027)])
def test_train_step_ragged_shapes_vs_oracle(dev, gemm_mode, B, N):
    """Shapes that are not multiples of the 128-row tiles (and a batch of one), seeded random weights, both GEMM engines,
    against the float64 oracle: loss, log-probabilities, head gradients, BatchNorm buffers, and the Adam update."""
    from oracle import oracle as orc
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.train import FlatAdam, cross_entropy, semseg_forward_train

    torch.manual_seed(77)
    net = PointNet2SemSeg(19, feature_dims=1)
    sd0 = orc.numpy_state_dict(net.state_dict())
    net = net.to(dev).train()
    opt = FlatAdam(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-4)
    rng = np.random.default_rng(B * 1000 + N)
    pts = syn.kitti_batch(B, N, config=5)
    target = rng.integers(0, 19, (B, N))
    starts = [rng.integers(0, n, B) for n in (N, 1024, 256, 64)]
    keep = rng.integers(0, 2, (B * N, 128)).astype(np.uint8)
    logp = semseg_forward_train(net, T(pts, dev), fps_starts=[T(s, dev) for s in starts], dropout_mask=T(keep, dev))
    loss = cross_entropy(logp, T(target, dev))
    opt.zero_grad()
    loss.backward()
    want = tor.semseg_train_step(sd0, pts, target, starts, keep)
    assert abs(loss.item() - want["loss"]) < 2e-4
    assert rl2(logp.detach().cpu().numpy(), want["logp"]) < 3e-4
    grads = _grads(net)
    for name in ("conv2.weight", "conv2.bias"):
        assert rl2(grads[name], want["grads"][name]) < 2e-3, (name, rl2(grads[name], want["grads"][name]))
    # sums of random-sign terms over the rows, gated by a ReLU mask: ONE mask flip (an activation within the engines' 1e-5
    # rounding difference of zero) moves such a sum by 1/sqrt(rows) ~ 2e-2 of its size
    flip_tol = 2e-3 if gemm_mode == "fp32" else 5e-2
    for name in ("bn1.weight", "bn1.bias"):
        assert rl2(grads[name], want["grads"][name]) < flip_tol, (name, rl2(grads[name], want["grads"][name]))
    # deeper gradients: every arg-max / ReLU flip on the way down moves whole gradient rows (few rows at sa4: 16 centroids per
    # cloud).  Exact-fp32 engine: within 5e-2 in L2; tensor-core engine (more flips): direction within 0.98 cosine.
    for name in ("conv1.weight", "fp1.mlp_convs.2.weight", "sa4.mlp_convs.0.weight", "sa1.mlp_convs.0.weight"):
        a_, b_ = grads[name].astype(np.float64).reshape(-1), np.asarray(want["grads"][name], np.float64).reshape(-1)
        if gemm_mode == "fp32":
            assert rl2(a_, b_) < 5e-2, (name, rl2(a_, b_))
        else:
            cos = float(a_ @ b_ / (np.linalg.norm(a_) * np.linalg.norm(b_)))
            assert cos > 0.98 and abs(np.linalg.norm(a_) / np.linalg.norm(b_) - 1) < 0.1, (name, cos)
    for name, buf in net.named_buffers():
        if name.endswith("num_batches_tracked"):
            assert int(buf) == 1
        else:
            ref = want["buffers"][name]
            assert np.abs(buf.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), name
    # the optimizer step on the flat buffer == torch.optim.Adam's formula applied to our gradients
    before = {n: p.detach().cpu().numpy().astype(np.float64) for n, p in net.named_parameters()}
    opt.step()
    for name in ("conv2.weight", "sa3.mlp_bns.1.bias", "fp2.mlp_convs.0.weight"):
        p_new, _, _ = tor.adam_step(before[name], grads[name], 0.0, 0.0, 1)
        got = dict(net.named_parameters())[name].detach().cpu().numpy()
        assert np.abs(got - p_new).max() < 2e-6, name


@pytest.mark.parametrize("tag,cls_name", [("ssg", "PointNet2ClsSsg"), ("msg", "PointNet2ClsMsg")])
def test_cls_nets_train_step_vs_reference(dev, golden, gemm_mode, tag, cls_name):
    """PointNet2ClsSsg / PointNet2ClsMsg in train() mode (group-all level, MSG blocks, fc head with two dropouts) against the
    reference's own autograd (tests/golden/train_cls_seeded.npz: forward, F.nll_loss, backward on 4 clouds x 1024 points)."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model import pointnet2 as ours

    g = golden("train_cls_seeded")
    torch.manual_seed(4242)
    net = getattr(ours, cls_name)().to(dev).train()
    xyz = T(syn.modelnet_batch(4, 1024, seed=4100), dev)
    masks = [T(g[f"{tag}.keep1"], dev), T(g[f"{tag}.keep2"], dev)]
    torch.manual_seed(9)                                   # the reference's two FPS start draws, same generator, same order
    out = net(xyz, dropout_masks=masks)
    logp = out[0] if isinstance(out, tuple) else out
    loss = torch.nn.functional.nll_loss(logp, T(g["target"], dev))
    net.zero_grad()
    loss.backward()
    assert abs(loss.item() - float(g[f"{tag}.loss"])) < 2e-4
    assert rl2(logp.detach().cpu().numpy(), g[f"{tag}.logp"]) < 2e-4
    grads = _grads(net)
    tight = 2e-3 if gemm_mode == "fp32" else 5e-2          # ReLU / arg-max flips (see the SA block test); 4 rows in the head
    bad = []
    for name, gr in grads.items():
        ref = g[f"{tag}.grad.{name}"]
        mine = gr.reshape(-1)
        if mine.size > 20000:
            mine = mine[::9]
        is_bias_before_bn = name.endswith(".bias") and ("convs" in name or "conv_blocks" in name or name in ("fc1.bias", "fc2.bias"))
        if is_bias_before_bn:
            assert np.abs(mine).max() < 1e-3, name          # analytically zero: BatchNorm follows
            continue
        a_, b_ = mine.astype(np.float64), ref.astype(np.float64)
        if np.linalg.norm(a_) < 1e-4 and np.linalg.norm(b_) < 1e-4:
            continue      # analytically zero as well: the beta of a level's last BatchNorm only shifts what the NEXT BatchNorm removes
        if name.startswith("fc3"):
            if not rl2(a_, b_) < tight:
                bad.append((name, "rl2", rl2(a_, b_)))
        else:
            cos = float(a_ @ b_ / max(np.linalg.norm(a_) * np.linalg.norm(b_), 1e-30))
            if not cos > (0.999 if gemm_mode == "fp32" else 0.98):
                bad.append((name, "cos", round(cos, 4), float(np.linalg.norm(a_)), float(np.linalg.norm(b_))))
    assert not bad, bad
    for name, buf in net.named_buffers():
        if not name.endswith("num_batches_tracked"):
            ref = g[f"{tag}.buffer.{name}"]
            assert np.abs(buf.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max()), name


def test_partseg_net_train_step_vs_reference(dev, golden, gemm_mode):
    """PointNet2PartSegSsg in train() mode: group-all level, feature propagation from a single coarse point, and a head with TWO
    outputs in the loss (log-probabilities and the pre-dropout features), against the reference's own autograd."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2PartSegSsg

    g = golden("train_cls_seeded")
    torch.manual_seed(4343)
    net = PointNet2PartSegSsg(50).to(dev).train()
    xyz = T(syn.modelnet_batch(2, 1024, seed=4200), dev)
    keep = T(np.unpackbits(g["part.keep_bits"], axis=1)[:, :128].astype(np.uint8), dev)
    torch.manual_seed(9)
    logp, feat = net(xyz, dropout_mask=keep)
    assert logp.shape == (2, 1024, 50) and feat.shape == (2, 128, 1024)
    target = T(g["part.target"].astype(np.int64), dev)
    loss = torch.nn.functional.nll_loss(logp.reshape(-1, 50), target.reshape(-1)) + 1e-3 * feat.pow(2).mean()
    net.zero_grad()
    loss.backward()
    assert abs(loss.item() - float(g["part.loss"])) < 2e-4
    assert rl2(logp.detach().cpu().numpy(), g["part.logp"]) < 2e-4
    assert rl2(feat.detach().cpu().numpy()[:, :, ::8], g["part.feat_sub"]) < (2e-4 if gemm_mode == "fp32" else 1e-3)
    bad = []
    for name, gr in _grads(net).items():
        ref = g[f"part.grad.{name}"].astype(np.float64)
        mine = gr.reshape(-1).astype(np.float64)
        if mine.size > 20000:
            mine = mine[::9]
        if name.endswith(".bias") and ("mlp_convs" in name or name == "conv1.bias"):
            assert np.abs(mine).max() < 1e-3, name         # analytically zero: BatchNorm follows (the reference holds round-off)
            continue
        if np.linalg.norm(mine) < 1e-4 and np.linalg.norm(ref) < 1e-4:
            continue                                       # analytically zero as well (the beta ahead of the next BatchNorm)
        cos = float(mine @ ref / max(np.linalg.norm(mine) * np.linalg.norm(ref), 1e-30))
        if not (cos > (0.999 if gemm_mode == "fp32" else 0.98) and abs(np.linalg.norm(mine) / np.linalg.norm(ref) - 1) < 0.05):
            bad.append((name, round(cos, 4), float(np.linalg.norm(mine)), float(np.linalg.norm(ref))))
    assert not bad, bad


def test_pointnet_seg_train_step_vs_reference(dev, golden, gemm_mode):
    """PointNetSeg(19, 4, feature_transform=True), the default model of the reference's training driver, one iteration as
    pcdseg.py:166-186 runs it (CrossEntropyLoss + 0.001 * feature_transform_reguliarzer) against the reference's autograd:
    input / feature transforms (STN, bmm with both gradients), BatchNorm without ReLU ahead of the global max, 1088-wide head."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet import PointNetSeg, feature_transform_reguliarzer

    g = golden("train_pointnet_seg_seeded")
    torch.manual_seed(4444)
    net = PointNetSeg(19, input_dims=4, feature_transform=True).to(dev).train()
    pts = T(syn.kitti_batch(8, 512, config=7), dev)       # 8 clouds: batch statistics of the STN's fc layers need more than 2 rows
    logits, trans_feat = net(pts)
    assert logits.shape == (8, 512, 19) and trans_feat.shape == (8, 64, 64)
    loss = torch.nn.CrossEntropyLoss()(logits.transpose(2, 1), T(g["target"].astype(np.int64), dev))
    loss = loss + feature_transform_reguliarzer(trans_feat) * 0.001
    net.zero_grad()
    loss.backward()
    bad = []
    if not abs(loss.item() - float(g["loss"])) < 3e-4:
        bad.append(("loss", loss.item(), float(g["loss"])))
    if not rl2(logits.detach().cpu().numpy(), g["logp"]) < (3e-4 if gemm_mode == "fp32" else 1e-3):
        bad.append(("logp", rl2(logits.detach().cpu().numpy(), g["logp"])))
    if not rl2(trans_feat.detach().cpu().numpy(), g["trans_feat"]) < (3e-4 if gemm_mode == "fp32" else 2e-3):
        bad.append(("trans_feat", rl2(trans_feat.detach().cpu().numpy(), g["trans_feat"])))
    for name, gr in _grads(net).items():
        ref = g["grad." + name].astype(np.float64)
        mine = gr.reshape(-1).astype(np.float64)
        if mine.size > 4096:
            mine = mine[::31]
        if np.linalg.norm(mine) < 1e-4 and np.linalg.norm(ref) < 1e-4:
            continue                                       # analytically zero (bias / beta ahead of a BatchNorm) or tiny
        if name.endswith(".bias") and np.abs(ref).max() < 1e-3 and np.abs(mine).max() < 1e-3:
            continue
        cos = float(mine @ ref / max(np.linalg.norm(mine) * np.linalg.norm(ref), 1e-30))
        if not (cos > (0.995 if gemm_mode == "fp32" else 0.97) and abs(np.linalg.norm(mine) / np.linalg.norm(ref) - 1) < 0.1):
            bad.append((name, round(cos, 4), float(np.linalg.norm(mine)), float(np.linalg.norm(ref))))
    assert not bad, bad
    for name, buf in net.named_buffers():
        if not name.endswith("num_batches_tracked"):
            ref = g["buffer." + name]
            assert np.abs(buf.cpu().numpy() - ref).max() < 2e-4 * max(1.0, np.abs(ref).max()), name


@pytest.mark.parametrize("groups,K,C", [(2, 1000, 70), (8, 512, 1024), (3, 257, 33)])
def test_tall_max_pool_with_argmax(dev, groups, K, C):
    """bn_act_max over all points of a cloud (K >= 256: one CTA per cloud and 32-channel slab): values and FIRST-maximum
    indices, with duplicated rows so that ties exist, and the gradient routing through them."""
    from pointnet12_b200 import ops

    rng = np.random.default_rng(K + C)
    y = rng.standard_normal((groups, K, C)).astype(np.float32)
    y[:, K // 2:K // 2 + 40] = y[:, 10:50]                   # exact duplicates: ties between row k and row k + K/2 - 10
    st = ops.BatchStats()
    st.scale, st.shift = T(rng.uniform(0.5, 1.5, C).astype(np.float32), dev), T(rng.standard_normal(C).astype(np.float32), dev)
    st.mean, st.invstd = st.shift, st.scale
    for relu in (True, False):
        pooled, am = ops.bn_act_max(T(y.reshape(groups * K, C), dev), st, K, relu=relu)
        z = y * st.scale.cpu().numpy() + st.shift.cpu().numpy()
        z32 = np.float32(y) * st.scale.cpu().numpy() + st.shift.cpu().numpy()
        if relu:
            z32 = np.maximum(z32, 0)
        got_am = am.cpu().numpy().astype(np.int64)
        got = pooled.cpu().numpy()
        # the kernel's own values (fused multiply-add) decide the arg-max; check it is A maximum and the FIRST one
        picked = np.take_along_axis(z32, got_am[:, None, :], 1)[:, 0, :]
        assert np.abs(got - z32.max(1)).max() < 1e-5 and np.abs(picked - got).max() < 1e-5
        first = (np.abs(z32 - got[:, None, :]) < 1e-6).argmax(1)
        assert (got_am <= first + 0).mean() > 0.999           # never later than the first row within rounding of the max


def test_graphed_step_pointnet_matches_eager(dev):
    """GraphedStep (generic CUDA-graph iteration) on PointNetSeg == the eager iteration with the same loss."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet import PointNetSeg, feature_transform_reguliarzer
    from pointnet12_b200.train import FlatAdam, GraphedStep, cross_entropy

    pts = T(syn.kitti_batch(4, 1024, config=7), dev)
    target = T(np.random.default_rng(2).integers(0, 19, (4, 1024)), dev)
    torch.manual_seed(5)
    a = PointNetSeg(19, 4, True).to(dev).train()
    b = PointNetSeg(19, 4, True).to(dev).train()
    b.load_state_dict(a.state_dict())

    def loss_fn(n, x, t):
        out, tf = n(x)
        return cross_entropy(out, t) + feature_transform_reguliarzer(tf) * 0.001

    opt_a, opt_b = FlatAdam(a.parameters(), lr=1e-3, weight_decay=1e-4), FlatAdam(b.parameters(), lr=1e-3, weight_decay=1e-4)
    runner = GraphedStep(b, opt_b, loss_fn)
    la, lb = [], []
    for _ in range(3):
        loss = loss_fn(a, pts, target)
        opt_a.zero_grad()
        loss.backward()
        opt_a.step()
        la.append(loss.item())
        lb.append(runner(pts, target).item())
    assert abs(la[0] - lb[0]) < 1e-5 and max(abs(x - y) / x for x, y in zip(la, lb)) < 2e-2, (la, lb)
    assert la[-1] < la[0] and int(b.bn1.num_batches_tracked) == 3


def test_pointnet_cls_train_step_vs_reference(dev, golden, gemm_mode):
    """PointNetCls(40, feature_transform=True): BatchNorm that follows a dropout instead of its linear layer
    (relu(bn2(dropout(fc2(x)))), pointnet.py:148), against the reference's autograd."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet import PointNetCls, feature_transform_reguliarzer

    g = golden("train_pointnet_seg_seeded")
    torch.manual_seed(4545)
    net = PointNetCls(40, feature_transform=True).to(dev).train()
    logp, tf = net(T(syn.modelnet_batch(8, 512, seed=4300), dev), dropout_mask=T(g["cls.keep"], dev))
    loss = torch.nn.functional.nll_loss(logp, T(g["cls.target"], dev)) + feature_transform_reguliarzer(tf) * 0.001
    net.zero_grad()
    loss.backward()
    bad = []
    if not abs(loss.item() - float(g["cls.loss"])) < 5e-4:
        bad.append(("loss", loss.item(), float(g["cls.loss"])))
    if not rl2(logp.detach().cpu().numpy(), g["cls.logp"]) < (5e-4 if gemm_mode == "fp32" else 2e-3):
        bad.append(("logp", rl2(logp.detach().cpu().numpy(), g["cls.logp"])))
    for name, gr in _grads(net).items():
        ref = g["cls.grad." + name].astype(np.float64)
        mine = gr.reshape(-1).astype(np.float64)
        if mine.size > 4096:
            mine = mine[::31]
        if np.linalg.norm(mine) < 1e-4 and np.linalg.norm(ref) < 1e-4:
            continue
        if name.endswith(".bias") and np.abs(ref).max() < 1e-3 and np.abs(mine).max() < 1e-3:
            continue
        cos = float(mine @ ref / max(np.linalg.norm(mine) * np.linalg.norm(ref), 1e-30))
        if not (cos > (0.995 if gemm_mode == "fp32" else 0.97) and abs(np.linalg.norm(mine) / np.linalg.norm(ref) - 1) < 0.1):
            bad.append((name, round(cos, 4), float(np.linalg.norm(mine)), float(np.linalg.norm(ref))))
    assert not bad, bad


def test_pointnet_densecls_train_step_vs_reference(dev, golden, gemm_mode):
    """PointNetDenseCls(16, 50): shared encoder with two heads (raw classification logits and per-point log-probabilities over a
    4944-channel concat), against the reference's autograd."""
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet import PointNetDenseCls, feature_transform_reguliarzer

    g = golden("train_pointnet_seg_seeded")
    torch.manual_seed(4646)
    net = PointNetDenseCls(16, 50).to(dev).train()
    cls_logits, seg_logp, tf = net(T(syn.modelnet_batch(8, 256, seed=4400), dev), T(g["dense.label"], dev),
                                   dropout_mask=T(g["dense.keep"], dev))
    assert cls_logits.shape == (8, 16) and seg_logp.shape == (8, 256, 50) and tf.shape == (8, 128, 128)
    loss = (torch.nn.functional.cross_entropy(cls_logits, T(g["dense.cls_target"], dev))
            + torch.nn.functional.nll_loss(seg_logp.reshape(-1, 50), T(g["dense.seg_target"].astype(np.int64), dev).reshape(-1))
            + feature_transform_reguliarzer(tf) * 0.001)
    net.zero_grad()
    loss.backward()
    bad = []
    if not abs(loss.item() - float(g["dense.loss"])) < 1e-3:
        bad.append(("loss", loss.item(), float(g["dense.loss"])))
    for key, got in (("dense.cls_logits", cls_logits), ("dense.seg_logp", seg_logp)):
        if not rl2(got.detach().cpu().numpy(), g[key]) < (1e-3 if gemm_mode == "fp32" else 5e-3):
            bad.append((key, rl2(got.detach().cpu().numpy(), g[key])))
    for name, gr in _grads(net).items():
        ref = g["dense.grad." + name].astype(np.float64)
        mine = gr.reshape(-1).astype(np.float64)
        if mine.size > 4096:
            mine = mine[::63]
        if np.linalg.norm(mine) < 1e-4 and np.linalg.norm(ref) < 1e-4:
            continue
        if name.endswith(".bias") and np.abs(ref).max() < 1e-3 and np.abs(mine).max() < 1e-3:
            continue
        cos = float(mine @ ref / max(np.linalg.norm(mine) * np.linalg.norm(ref), 1e-30))
        if not (cos > (0.99 if gemm_mode == "fp32" else 0.95) and abs(np.linalg.norm(mine) / np.linalg.norm(ref) - 1) < 0.15):
            bad.append((name, round(cos, 4), float(np.linalg.norm(mine)), float(np.linalg.norm(ref))))
    assert not bad, bad
