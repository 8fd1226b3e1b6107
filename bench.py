"""bench.py -- PointNet++ SSG semantic-segmentation forward, points/sec (BASELINE.json metric, config C2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One step = one eval forward of PointNet2SemSeg (the reference's shipped pointnet2-inview checkpoint) over a
batch of 8 synthetic KITTI-shaped clouds of 24 000 points (xyz + reflectance) per GPU.  Clouds are
independent, so ranks shard the work with no data-path collective (weak scaling: 8 clouds per GPU).

Our arm prints one JSON line with
  value         whole-job points/sec, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e           the same through the public nn.Module call with pinned HOST input and HOST output
                (H2D of the batch and D2H of the [B,N,19] log-probs inside the timed region)
  roofline      the dominant kernel (level-1 farthest-point sampling, pn_fps_f32 at N=24000): algorithmic
                bytes B*npoint*N*16 per launch / mean launch duration measured with CUDA events inside the
                timed steps, against the measured HBM copy bandwidth of MEASURED_PEAKS.json
  cpu_baseline  the CPU oracle (a C/OpenMP port of the reference algorithm, oracle/) on the same batch
`--impl reference` times that CPU oracle port as the reference arm (the reference itself is pure Python and
cannot travel to the GPU box; the port is ~10x faster than the reference's own PyTorch-CPU path).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CKPT = os.path.join(ROOT, "tests", "golden", "pointnet2-inview-0.55884-0001.pth")
BATCH, NPOINTS, CLASSES = 8, 24000, 19
LEVEL_N = (24000, 1024, 256, 64)
METRIC = "pointnet2_semseg_forward_points_per_sec"
FPS_DRAM_BYTES_PER_LAUNCH = 2337024   # ncu: 2.34 MB read + 0 written per level-1 launch (the cloud lives in registers)
FPS_ENTRY_POINTS = ["pn_fps_f32", "pn_fps_progress_f32"]   # the same kernel with / without the progress feed
WORKLOAD = "C2: PointNet2SemSeg(19, feature_dims=1) eval forward, pointnet2-inview checkpoint, 8 synthetic KITTI-shaped clouds x 24000 points per GPU"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle leg (our arm only)")
    ap.add_argument("--pdl", action="store_true", help="programmatic dependent launch of the critical-path kernels")
    ap.add_argument("--precision", default="bf16x3", choices=["bf16x3", "bf16", "fp32"],
                    help="bf16x3 (default, fp32 parity), bf16 (single-pass, stated tolerance) or fp32 (CUDA cores)")
    ap.add_argument("--workload", default="c2", choices=["c2", "train", "preprocess"],
                    help="c2 (default): the headline eval forward; train: config C5, one training iteration per step "
                         "(tools/bench_train.py; same JSON contract); preprocess: raw scans -> network input (tools/bench_preprocess.py)")
    ap.add_argument("--mode", default="graph", choices=["graph", "eager"],
                    help="graph: one CUDA-graph replay per step (default); eager: ~40 launches per step")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons of one GPU while the timed region runs.

    The timed region is ~20 ms, shorter than one `nvidia-smi -lms` period (and than nvidia-smi's start-up on an 8-GPU
    box), so the samples come from NVML directly -- the library nvidia-smi itself reads -- every 2 ms in a thread:
    clocks.sm, clocks.max.sm and the clocks_event_reasons bits hw_slowdown / hw_thermal_slowdown / sw_thermal_slowdown /
    sw_power_cap of the recipe's clocks line.  nvidia-smi is the fallback when NVML cannot be loaded."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc, self.nvml, self.stop = index, [], None, None, threading.Event()
        self.source = None

    def __enter__(self):
        try:
            import pynvml

            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml, self.source = pynvml, "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return self
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _poll_nvml(self):
        n = self.nvml
        bits = [("hw_slowdown", n.nvmlClocksEventReasonHwSlowdown if hasattr(n, "nvmlClocksEventReasonHwSlowdown")
                 else n.nvmlClocksThrottleReasonHwSlowdown),
                ("hw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonHwThermalSlowdown",
                                                getattr(n, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40))),
                ("sw_thermal_slowdown", getattr(n, "nvmlClocksEventReasonSwThermalSlowdown",
                                                getattr(n, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20))),
                ("sw_power_cap", getattr(n, "nvmlClocksEventReasonSwPowerCap",
                                         getattr(n, "nvmlClocksThrottleReasonSwPowerCap", 0x4)))]
        reasons_fn = getattr(n, "nvmlDeviceGetCurrentClocksEventReasons", None) or n.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = n.nvmlDeviceGetMaxClockInfo(self.handle, n.NVML_CLOCK_SM)
        while not self.stop.is_set():
            try:
                sm = n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)
                mask = reasons_fn(self.handle)
                self.rows.append([str(sm), str(mx)] + ["Active" if mask & b else "Not Active" for _, b in bits])
            except Exception:
                pass
            time.sleep(0.002)

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *exc):
        if self.nvml is not None:
            self.stop.set()
            self.thread.join(timeout=2)
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
        elif self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 6 for n, v in zip(names, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


def fps_starts(torch, batch):
    """The four FPS start-index draws of one forward, exactly as the modules draw them (pointnet_util.py:75)."""
    return [torch.randint(0, n, (batch,), dtype=torch.long) for n in LEVEL_N]


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_forward_timer(batch):
    import torch

    from oracle import oracle as orc
    from pointnet12_b200 import synthetic as syn

    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm uses all host cores explicitly
    orc.set_num_threads(os.cpu_count() or 1)
    sd = orc.numpy_state_dict(torch.load(CKPT, map_location="cpu"))
    pts = syn.kitti_batch(batch, NPOINTS, config=2)
    torch.manual_seed(0)
    starts = [s.numpy() for s in fps_starts(torch, batch)]

    def step():
        t = time.perf_counter()
        orc.pointnet2_semseg(sd, pts, starts)
        return time.perf_counter() - t

    return step, orc.num_threads()


def run_reference(args):
    """Reference arm: the CPU port of the reference algorithm on the host cores, same config/metric/unit."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    step, threads = oracle_forward_timer(BATCH)
    for _ in range(max(1, min(args.warmup, 2))):
        step()
    times = [step() for _ in range(args.steps)]
    total = sum(times)
    value = BATCH * NPOINTS * args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "batch_per_step": BATCH, "points_per_cloud": NPOINTS},
        "cpu_baseline": {"value": value, "unit": "points/s", "cores": threads, "kind": "port",
                         "sample": f"full step: {BATCH} clouds x {NPOINTS} points per step, {args.steps} steps, "
                                   f"C/OpenMP oracle (oracle/pn_oracle.c) on {os.cpu_count()} host CPUs"},
        "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from pointnet12_b200 import _native as nv
    from pointnet12_b200 import ops
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.utils import load_pointnet
    from pointnet12_b200.runtime import GraphedSemSeg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    if args.pdl:
        ops.set_pdl(True)
    ops.set_mlp_mode(args.precision)
    net = load_pointnet("pointnet2", CLASSES, CKPT, device=dev)
    # four different batches per rank, rotated, so consecutive steps never see the same clouds
    host_batches = [torch.from_numpy(syn.kitti_batch(BATCH, NPOINTS, config=2, first=(rank * 4 + i) * BATCH)).pin_memory()
                    for i in range(4)]
    dev_batches = [h.to(dev) for h in host_batches]
    host_out = torch.empty((BATCH, NPOINTS, CLASSES), dtype=torch.float32).pin_memory()
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    runner = GraphedSemSeg(net) if args.mode == "graph" else None

    def step_eager(i):
        with torch.no_grad():
            return net(dev_batches[i % 4])

    def step_resident(i):
        if runner is not None:
            return runner(dev_batches[i % 4])
        return step_eager(i)

    def step_e2e(i):
        if runner is not None:
            runner(host_batches[i % 4], to_host=True)    # D2H copies are graph nodes; result in the runner's pinned buffer
            return
        with torch.no_grad():
            x = host_batches[i % 4].to(dev, non_blocking=True)
            net(x, host_out=host_out)

    def timed(step_fn, steps):
        """Sum of per-step CUDA-event durations; L2 is flushed (untimed) between steps."""
        evs = []
        for i in range(steps):
            flush.fill_(i & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step_fn(i)
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        per_step = [a.elapsed_time(b) for a, b in evs]
        timed.last = per_step
        return sum(per_step)      # ms

    torch.manual_seed(1234 + rank)
    for i in range(max(args.warmup, 3)):
        step_resident(i)
        step_e2e(i)
    barrier()

    # ---- timed region 1: inputs resident in HBM
    if runner is None:
        nv.time_entry_points(FPS_ENTRY_POINTS)   # eager: the dominant kernel's launches carry their own events
    launches0 = nv.launch_count
    with ClockSampler(local) as clocks:
        ms = timed(step_resident, args.steps)
    barrier()
    spread = sorted(timed.last)
    launches = nv.launch_count - launches0
    if runner is None:
        fps_records = sum(nv.time_entry_points(None).values(), [])
    else:
        # a graph node cannot be bracketed by events: time the identical kernel in an eager pass of the same steps
        nv.time_entry_points(FPS_ENTRY_POINTS)
        l0 = nv.launch_count
        timed(step_eager, args.steps)
        launches = nv.launch_count - l0            # kernels per step x steps = nodes the graph replays
        fps_records = sum(nv.time_entry_points(None).values(), [])
        barrier()

    # ---- timed region 2: end to end from pinned host memory and back
    barrier()
    ms_e2e = timed(step_e2e, args.steps)
    barrier()

    t = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    points = world * BATCH * NPOINTS * args.steps

    if rank == 0:
        peak, peak_src = measured_peaks()
        l1 = [a.elapsed_time(b) for a, b, tag in fps_records if tag == (BATCH, NPOINTS, 1024)]
        fps_ms = sum(l1) / len(l1)
        fps_bytes = BATCH * 1024 * NPOINTS * 16
        achieved = fps_bytes / (fps_ms * 1e-3) / 1e9
        all_fps_ms = sum(a.elapsed_time(b) for a, b, _ in fps_records) / args.steps
        line = {
            "metric": METRIC, "value": points / (ms * 1e-3), "unit": "points/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16x3": "f32 (MLP products as split bf16 hi/lo on tensor cores)", "bf16": "bf16 (fp32 accumulation)",
                      "fp32": "f32"}[ops.mlp_precision()],
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": BATCH, "points_per_cloud": NPOINTS,
                       "precision": {"bf16x3": "bf16x3: tcgen05 tensor cores, 3-pass split bf16 with fp32 accumulation (fp32 "
                                               "parity, ~1e-5 relative)",
                                     "bf16": "bf16: tcgen05 tensor cores, single pass (max |delta log-prob| 0.38, 99.6 % equal "
                                             "labels at this config)",
                                     "fp32": "fp32 FMA on CUDA cores"}[ops.mlp_precision()],
                       "launch": ("one CUDA-graph replay per step (5 internal streams forked/joined inside the graph)"
                                  if runner is not None else "eager launches on 5 internal streams"),
                       "l2": "512 MiB written between timed steps",
                       "fps_start": "torch.randint on the CPU generator per level, as the reference draws it"},
            "e2e": {"value": points / (ms_e2e * 1e-3), "unit": "points/s", "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": BATCH * 4 * NPOINTS * 4 + 4 * BATCH * 8,
                    "d2h_bytes_per_step": BATCH * NPOINTS * CLASSES * 4},
            "gpu_launches": launches,
            "step_ms": {"min": spread[0], "median": spread[len(spread) // 2], "max": spread[-1]},
            "roofline": {"kernel": "fps_async_kernel<4,24,true> (pn_fps_progress_f32, level 1: N=24000 -> 1024 centroids)",
                         "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": FPS_DRAM_BYTES_PER_LAUNCH,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture "
                                           "(profiles/r01_v11_ncu_full_summary.csv)",
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": fps_bytes,
                         "launch_ms": fps_ms, "share_of_step": fps_ms / (ms / args.steps),
                         "all_fps_levels_ms_per_step": all_fps_ms,
                         "timing": ("CUDA events around the launch, eager pass over the same steps right after the timed "
                                    "graph replays" if runner is not None else "CUDA events around the launch inside the timed steps"),
                         "note": "coordinates and running distances are register-resident, so DRAM traffic is ~0; "
                                 "the figure is effective bandwidth on SURVEY 8(d)'s algorithmic bytes"},
            "clocks": clocks.summary(),
        }
        if not args.no_cpu_baseline:
            step, threads = oracle_forward_timer(BATCH)
            step()
            times = [step() for _ in range(5)]
            line["cpu_baseline"] = {"value": BATCH * NPOINTS * len(times) / sum(times), "unit": "points/s",
                                    "cores": threads, "kind": "port",
                                    "sample": f"5 forwards of the same {BATCH} x {NPOINTS} batch after 1 warm-up, C/OpenMP "
                                              f"oracle on {os.cpu_count()} host CPUs"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_train_oracle(B=2, N=2048):
    """The float64 numpy oracle of one training iteration (oracle/train_oracle.py) on a bounded sample: points/s."""
    import numpy as np
    import torch

    from oracle import oracle as orc
    from oracle import train_oracle as tor
    from pointnet12_b200 import synthetic as syn
    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg

    torch.manual_seed(1234)
    sd = orc.numpy_state_dict(PointNet2SemSeg(19, feature_dims=1).state_dict())
    pts = syn.kitti_batch(B, N, config=5)
    rng = np.random.default_rng(0)
    target = rng.integers(0, 19, (B, N))
    starts = [rng.integers(0, n, B) for n in (N, 1024, 256, 64)]
    keep = rng.integers(0, 2, (B * N, 128))
    t0 = time.perf_counter()
    tor.semseg_train_step(sd, pts, target, starts, keep)
    dt = time.perf_counter() - t0
    return {"value": B * N / dt, "unit": "points/s", "cores": orc.num_threads(), "kind": "port",
            "sample": f"one training iteration (forward, loss, backward) of {B} clouds x {N} points, numpy float64 oracle "
                      f"(BLAS threads) + C/OpenMP geometry, {dt:.2f} s"}


def cpu_preprocess_oracle(scan, npoints, scans):
    """CPU leg of --workload preprocess: the numpy restatement of the reference's loader path on ONE raw scan."""
    import numpy as np

    from oracle import preprocess_oracle as por
    from pointnet12_b200 import synthetic as syn

    t0 = time.perf_counter()
    k, _ = por.scan_filter(*scan, syn.SEMANTIC_KITTI_LEARNING_MAP)
    rng = np.random.default_rng(0)
    por.scan_sample(*scan, syn.SEMANTIC_KITTI_LEARNING_MAP, npoints, rng.integers(0, len(k), npoints),
                    noise=np.zeros((len(k), 4), np.float32))
    ms = (time.perf_counter() - t0) * 1e3
    return {"value": scan[0].shape[0] / (ms * 1e-3), "unit": "points/s", "cores": 1, "kind": "port",
            "sample": f"one raw scan of {scan[0].shape[0]} points through oracle/preprocess_oracle.py (numpy), {ms:.1f} ms; "
                      f"a batch of {scans} scans = {ms * scans:.0f} ms"}


def main():
    args = parse()
    if args.workload in ("train", "preprocess") and args.impl == "ours":
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        if args.workload == "preprocess":
            import bench_preprocess

            bench_preprocess.main([], cpu_baseline=None if args.no_cpu_baseline else cpu_preprocess_oracle)
            return
        import bench_train

        bench_train.main(["--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--eager"] if args.mode == "eager" else [])
                         + (["--no-cpu-baseline"] if args.no_cpu_baseline else []),
                         cpu_baseline=None if args.no_cpu_baseline else cpu_train_oracle)
        return
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
