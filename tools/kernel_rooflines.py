"""Stand-alone timing of every kernel group of the PointNet2SemSeg forward (config C2) with its roofline.

The captured forward overlaps its kernels on six streams, so a kernel's duration inside the step says little about
the kernel itself; here every group runs ALONE on the GPU on the real intermediate tensors of one forward
(CUDA events, 512 MiB written before every repetition = cold L2, median of `reps`), and is reported against
  * the measured HBM copy bandwidth for the geometry kernels, on the ALGORITHMIC bytes of SURVEY.md section 8(d):
      farthest-point sampling  B * npoint * N * 16        ball query  B * S * N * 12 + B * S * nsample * 8
      3-NN search              B * N * S * 12
    (these kernels prune -- register-resident clouds, grid buckets, block bounds -- so the figure is an EFFECTIVE bandwidth
    and may exceed the peak; `traffic` is the DRAM traffic ncu measured, profiles/r02_ncu_traffic.json);
  * the measured dense bf16 tensor throughput for the fused chains, on the USEFUL flops 2 * rows * sum(cin * cout) with the
    true channel counts (the fp32-parity mode issues three bf16 products per useful one: `issued_frac` = 3 x).

    python tools/kernel_rooflines.py            (prints one JSON object; bench.py embeds the same list as `roofline_all`)
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), float(p["bf16_tflops"]), "measured (MEASURED_PEAKS.json; burst figures: kernels timed alone)"
    return 6650.0, 1500.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f)
    return {}


def _chain_flops(rows, convs):
    return 2.0 * rows * sum(c.weight.shape[0] * c.weight.shape[1] for c in convs)


def measure(net, points, starts, reps=5, fps1_config=None):
    """net: PointNet2SemSeg (eval) on the device, points [B,4,N] on the device, starts: the four FPS start vectors (device).
    fps1_config: launch shape (cluster, threads, exchange) of the level-1 sampling the measured step really uses (the runner
    with batches in flight samples on 2-3 CTAs per cloud); the one-batch-at-a-time shape is reported next to it."""
    from pointnet12_b200 import ops

    n = net.module if hasattr(net, "module") else net
    dev = points.device
    B, _, N = points.shape
    pm = points.permute(0, 2, 1)
    x0, f0 = pm[:, :, :3], pm[:, :, 3:]
    sa = [n.sa1, n.sa2, n.sa3, n.sa4]
    fp = [n.fp1, n.fp2, n.fp3, n.fp4]
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
    hbm, tflops, peak_src = peaks()
    traffic = ncu_traffic()
    out = []

    def timed(fn):
        res = fn()
        torch.cuda.synchronize()
        ts = []
        for i in range(reps):
            flush.fill_(i)                      # cold L2; also lets the host run ahead of the device
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            res = fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts), res

    def hbm_row(name, kernel, ms, nbytes, key):
        ach = nbytes / (ms * 1e-3) / 1e9
        out.append({"name": name, "kernel": kernel, "bound": "hbm", "launch_ms": ms, "algorithmic_bytes": int(nbytes),
                    "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic.get(key)})

    def tc_row(name, kernel, ms, flops, key):
        ach = flops / (ms * 1e-3) / 1e12
        passes = 3 if ops.mlp_precision() == "bf16x3" else 1
        out.append({"name": name, "kernel": kernel, "bound": "tensor", "launch_ms": ms, "useful_flops": int(flops),
                    "achieved": ach, "peak": tflops, "unit": "TFLOP/s", "frac": ach / tflops,
                    "issued_frac": passes * ach / tflops, "traffic": traffic.get(key)})

    with torch.no_grad():
        xs, fs, balls = [x0], [f0], []
        # ---- level 1: sampling, buckets, ball query, chain
        S, K, r = sa[0].npoint, sa[0].nsample, sa[0].radius
        auto1 = tuple(ops.fps1_config() or (0, 0, 0))          # the one-batch-at-a-time shape (automatic unless PN12_FPS1 is set)
        cfg1 = tuple(fps1_config) if fps1_config is not None else auto1
        ms, fps1 = timed(lambda: ops.fps(x0, S, starts[0], config=cfg1))
        hbm_row("fps level 1", "fps_async_kernel (pn_fps_f32)", ms, B * S * N * 16, "fps1")
        ctas, _ = ops.fps_launch_info(B, N, S, cfg1)
        # the kernel keeps the cloud in registers: what bounds it is the FP32 pipe of the SMs it occupies -- 12 lane operations
        # per point and iteration (3 sub, 3 mul, 2 add, min, compare, 2 selects), nothing fusable without changing the
        # reference's rounding
        out[-1].update({"launch_shape": {"cluster": cfg1[0], "threads": cfg1[1], "exchange": cfg1[2]}, "sms_occupied": ctas,
                        "fp32_lane_ops": int(B) * S * N * 12,
                        "fp32_pipe_frac_on_its_sms": B * S * N * 12 / (ms * 1e-3 * ctas * 128 * 1.965e9)})
        if cfg1 != auto1:
            ms_l, _ = timed(lambda: ops.fps(x0, S, starts[0], config=auto1))
            hbm_row("fps level 1, one batch at a time (latency-optimal shape, GraphedSemSeg depth 1)", "fps_async_kernel (pn_fps_f32)",
                    ms_l, B * S * N * 16, "fps1")
            ctas_l, _ = ops.fps_launch_info(B, N, S, auto1)
            out[-1].update({"sms_occupied": ctas_l, "fp32_pipe_frac_on_its_sms": B * S * N * 12 / (ms_l * 1e-3 * ctas_l * 128 * 1.965e9)})
        x1 = ops.index_points(x0, fps1)
        ms, grid1 = timed(lambda: ops.ball_grid(x0, r))
        hbm_row("ball-query buckets level 1", "ball_grid_build_kernel (pn_ball_grid_build_f32)", ms, B * N * 16 * 2, "grid1")
        ms, ball1 = timed(lambda: ops.ball_query(r, K, x0, x1, grid=grid1))
        hbm_row("ball query level 1 (stand-alone, after sampling)", "ball_query_grid_kernel (pn_ball_query_grid_f32)", ms,
                B * S * N * 12 + B * S * K * 8, "ball1")
        xs.append(x1)
        balls.append(ball1)
        ms, f1 = timed(lambda: sa[0].features(x0, f0, x1, ball1))
        tc_row("sa1 chain (gather + 3 layers + max)", "mlp_tc_res_kernel<SA,4> (pn_sa_mlp_bf16x3)", ms,
               _chain_flops(B * S * K, sa[0].mlp_convs), "sa1")
        fs.append(f1)
        # ---- levels 2..4
        for i in (1, 2, 3):
            S, K, r, Nl = sa[i].npoint, sa[i].nsample, sa[i].radius, xs[i].shape[1]
            ms, fi = timed(lambda: ops.fps(xs[i], S, starts[i]))
            hbm_row(f"fps level {i + 1}", "fps_kernel (pn_fps_f32)", ms, B * S * Nl * 16, f"fps{i + 1}")
            nx = ops.index_points(xs[i], fi)
            ms, bi = timed(lambda: ops.ball_query(r, K, xs[i], nx))
            hbm_row(f"ball query level {i + 1}", "ball_query_kernel (pn_ball_query_f32)", ms, B * S * Nl * 12 + B * S * K * 8,
                    f"ball{i + 1}")
            xs.append(nx)
            balls.append(bi)
            ms, fi2 = timed(lambda: sa[i].features(xs[i], fs[i], nx, bi))
            kern = "mlp_tc_res_kernel<SA,8>" if i == 1 else "mlp_tc_kernel<SA,256>"
            tc_row(f"sa{i + 1} chain", f"{kern} (pn_sa_mlp_bf16x3)", ms, _chain_flops(B * S * K, sa[i].mlp_convs), f"sa{i + 1}")
            fs.append(fi2)
        # ---- 3-NN searches
        nns = [None] * 4
        ms, nns[0] = timed(lambda: fp[0].geometry(x0, xs[1], order=grid1))
        hbm_row("3-NN for fp1 (block search incl. block build)", "nn_blocks_build_kernel + three_nn_blocks_kernel (pn_three_nn_blocks_f32)",
                ms, B * N * xs[1].shape[1] * 12, "nn1")
        for i in (1, 2, 3):
            ms, nns[i] = timed(lambda: fp[i].geometry(xs[i], xs[i + 1]))
            hbm_row(f"3-NN for fp{i + 1}", "three_nn_kernel (pn_three_nn_f32)", ms, B * xs[i].shape[1] * xs[i + 1].shape[1] * 12, f"nn{i + 1}")
        # ---- feature propagation
        up = fs[4]
        for i in (3, 2, 1):
            ms, up = timed(lambda: fp[i].features(fs[i], up, *nns[i]))
            tc_row(f"fp{i + 1} chain (interpolate + concat + layers)", "mlp_tc_kernel<FP,512> (pn_fp_mlp_bf16x3)", ms,
                   _chain_flops(B * xs[i].shape[1], fp[i].mlp_convs), f"fp{i + 1}")
        head = (n._head, [n.conv1, n.conv2], [n.bn1, None], [True, False], ops.OUT_LOG_SOFTMAX)
        ms, logp = timed(lambda: fp[0].features(None, up, *nns[0], head=head))
        # fp1's first layer runs on the coarse points (folded; in the captured forward it rides in fp2's launch)
        flops = (_chain_flops(B * xs[1].shape[1], fp[0].mlp_convs[:1]) + _chain_flops(B * N, list(fp[0].mlp_convs[1:]) + [n.conv1, n.conv2]))
        tc_row("fp1 + segmentation head chain (interpolate + 4 layers + log_softmax; first layer folded to the coarse level)",
               "mlp_tc_res_kernel<FP,8> (pn_fp_mlp_bf16x3) + the folded layer (pn_mlp_rows_bf16x3)", ms, flops, "fp1")
    return {"peak_source": peak_src, "timing": f"each group alone on the GPU, cold L2, CUDA events, median of {reps}",
            "kernels": out, "sum_ms": sum(k["launch_ms"] for k in out)}


if __name__ == "__main__":
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.utils import load_pointnet

    dev = torch.device("cuda", 0)
    net = load_pointnet("pointnet2", 19, os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth"), device=dev)
    x = torch.from_numpy(syn.kitti_batch(8, 24000, config=2)).to(dev)
    torch.manual_seed(0)
    st = [torch.randint(0, m, (8,), dtype=torch.long).to(dev) for m in (24000, 1024, 256, 64)]
    shape = tuple(int(t) for t in sys.argv[sys.argv.index("--fps1") + 1].split(",")) if "--fps1" in sys.argv else None
    print(json.dumps(measure(net, x, st, fps1_config=shape), indent=1))
