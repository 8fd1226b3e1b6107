"""Data-parallel plumbing: one process per GPU, clouds sharded across ranks, no data-path collective.

Clouds are independent in the eval forward (BatchNorm uses running statistics), so the reference's
torch.nn.DataParallel scatter/gather (model/utils.py:22, pcdseg.py:141) becomes: every rank takes a contiguous
slice of the global batch, runs the same replicated weights, keeps its outputs.  torch.distributed (NCCL on
GPUs, gloo in the CPU tests) is only used for the rendezvous, for timing reductions and for optional
gathers of results.
"""
from __future__ import annotations

import os
from typing import List, Tuple

import torch
import torch.distributed as dist


def env_world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment; (0, 1, 0) when run directly."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def _parse_cpulist(text: str) -> List[int]:
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.extend(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_node(index: int) -> int:
    """NUMA node of a GPU's PCIe slot from sysfs, or -1 (one node, a virtualised topology, no sysfs)."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            return int(f.read().strip())
    except Exception:   # noqa: BLE001
        return -1


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin the calling process to the CPUs of its GPU's NUMA node, BEFORE it allocates pinned host buffers: the staging
    memory of the end-to-end path (3 MB in, 14.6 MB out per batch and GPU) then sits next to the PCIe root the copies cross
    instead of on whichever node the launcher started the process on (round 1: eight ranks on node 0, 73 % of linear end
    to end at 8 GPUs).  Returns what was done, for the benchmark record; a no-op where the topology does not say."""
    info = {"gpu": local_rank, "numa_node": gpu_numa_node(local_rank), "cpus": None, "bound": False}
    node = info["numa_node"]
    if node < 0 or not hasattr(os, "sched_setaffinity"):
        return info
    try:
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["cpus"], info["bound"] = len(allowed), True
    except Exception:   # noqa: BLE001
        pass
    return info


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of n_items for this rank; the first n_items % world ranks get one extra."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(points: torch.Tensor, rank: int, world: int) -> torch.Tensor:
    """This rank's clouds of a global [B, C, N] batch (a view)."""
    lo, hi = shard_range(points.shape[0], rank, world)
    return points[lo:hi]


def max_over_ranks(values: List[float], device=None) -> List[float]:
    """Element-wise MAX over ranks (device timings are reported as the slowest rank's)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(values)
    t = torch.tensor(values, dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_labels(labels: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather per-rank [B_local, N] label maps into [global_batch, N] on every rank (uneven shards allowed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return labels
    world, rank = dist.get_world_size(), dist.get_rank()
    biggest = max(shard_range(global_batch, r, world)[1] - shard_range(global_batch, r, world)[0] for r in range(world))
    pad = torch.zeros((biggest,) + tuple(labels.shape[1:]), dtype=labels.dtype, device=labels.device)
    pad[:labels.shape[0]] = labels
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    out = []
    for r, p in enumerate(parts):
        lo, hi = shard_range(global_batch, r, world)
        out.append(p[:hi - lo])
    return torch.cat(out, 0)
