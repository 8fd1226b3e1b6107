"""Per-kernel device timings (CUDA events) for the sampling / grouping primitives and the full forward.

    python tools/microbench.py [--quick]

Config C3 of BASELINE.json: FPS + ball query + grouping over N = 4k..120k, npoint 1024/256/64, nsample 32,
reported as effective GB/s on the algorithmic bytes of SURVEY.md section 8(d):
  FPS        B * npoint * N * 16 bytes
  ball query B * S * N * 12 + B * S * nsample * 8 bytes
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from pointnet12_b200 import ops, synthetic as syn  # noqa: E402


def time_ms(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    quick = "--quick" in sys.argv
    dev = torch.device("cuda", 0)
    out = []
    Ns = [4096, 24000] if quick else [4096, 8192, 16384, 24000, 32768, 65536, 120000]
    for B in ([] if '--only-sweeps' in sys.argv else [8] if quick else [1, 8]):
        for N in Ns:
            pts = torch.from_numpy(syn.kitti_batch(B, N, config=3)).to(dev)
            xyz = pts.permute(0, 2, 1)[:, :, :3]
            feat = pts.permute(0, 2, 1)[:, :, 3:]
            for npoint, radius in ((1024, 0.1), (256, 0.2), (64, 0.4)):
                start = torch.zeros(B, dtype=torch.long, device=dev)
                t_fps = time_ms(lambda: ops.fps(xyz, npoint, start))
                fps_idx = ops.fps(xyz, npoint, start)
                new_xyz = ops.index_points(xyz, fps_idx)
                t_ball = time_ms(lambda: ops.ball_query(radius, 32, xyz, new_xyz))
                idx = ops.ball_query(radius, 32, xyz, new_xyz)
                t_grp = time_ms(lambda: ops.group(xyz, feat, new_xyz, idx, False))
                fps_b = B * npoint * N * 16
                ball_b = B * npoint * N * 12 + B * npoint * 32 * 8
                rec = dict(B=B, N=N, npoint=npoint, radius=radius, fps_ms=round(t_fps, 4), ball_ms=round(t_ball, 4),
                           group_ms=round(t_grp, 4), fps_gbs=round(fps_b / t_fps / 1e6, 1),
                           ball_gbs=round(ball_b / t_ball / 1e6, 1))
                print(json.dumps(rec), flush=True)
                out.append(rec)
    if "--fps-sweep" in sys.argv:
        B, N, npoint = 8, 24000, 1024
        pts = torch.from_numpy(syn.kitti_batch(B, N, config=2)).to(dev)
        xyz = pts.permute(0, 2, 1)[:, :, :3]
        start = torch.zeros(B, dtype=torch.long, device=dev)
        for ex, cl, th in [(1, 8, 256), (1, 8, 512), (2, 4, 128), (2, 4, 256), (2, 8, 64), (2, 8, 128), (2, 8, 256),
                           (2, 16, 64), (2, 16, 128)]:
            if True:
                try:
                    ops.fps_set_config(cl, th, ex)
                    t = time_ms(lambda: ops.fps(xyz, npoint, start))
                    print(json.dumps(dict(sweep="fps", exchange=ex, cluster=cl, threads=th, ms=round(t, 4),
                                          us_per_iter=round(t * 1000 / npoint, 3),
                                          gbs=round(B * npoint * N * 16 / t / 1e6, 1))), flush=True)
                except RuntimeError as e:
                    print(json.dumps(dict(sweep="fps", cluster=cl, threads=th, error=str(e)[:100])), flush=True)
                finally:
                    ops.fps_set_config(0, 0, 0)
    if "--ball-sweep" in sys.argv:
        for B, N, npoint, radius in [(8, 24000, 1024, 0.1), (1, 120000, 1024, 0.1), (8, 8192, 1024, 0.1)]:
            pts = torch.from_numpy(syn.kitti_batch(B, N, config=2)).to(dev)
            xyz = pts.permute(0, 2, 1)[:, :, :3]
            start = torch.zeros(B, dtype=torch.long, device=dev)
            new_xyz = ops.index_points(xyz, ops.fps(xyz, npoint, start))
            t_build = time_ms(lambda: ops.ball_grid(xyz, radius))
            grid = ops.ball_grid(xyz, radius)
            for thr in (512, 1024, 2048, 4096, 8192):
                t = time_ms(lambda: ops.ball_query(radius, 32, xyz, new_xyz, grid=grid, method="grid", threshold=thr))
                print(json.dumps(dict(sweep="ball-threshold", B=B, N=N, S=npoint, threshold=thr, ms=round(t, 4))), flush=True)
            for method in ("scan", "grid", "grid-cells", "grid-scan"):
                t = time_ms(lambda: ops.ball_query(radius, 32, xyz, new_xyz, grid=None if method == "scan" else grid,
                                                   method=method))
                print(json.dumps(dict(sweep="ball", B=B, N=N, S=npoint, method=method, ms=round(t, 4),
                                      build_ms=round(t_build, 4),
                                      gbs=round((B * npoint * N * 12 + B * npoint * 256) / t / 1e6, 1))), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
