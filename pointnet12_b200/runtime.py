"""CUDA-graph replay of a network forward (launch-bound inner loop -> one graph launch per batch).

A PointNet2SemSeg forward is ~40 kernel launches spread over three streams; captured once per input shape it
replays as a single graph launch, with the fork/join between the streams preserved as graph dependencies.
Only device memory, streams and graphs come from PyTorch; every node of the graph is one of our kernels.

    runner = GraphedSemSeg(net)            # net: PointNet2SemSeg or the load_pointnet wrapper, in eval mode
    logp = runner(points)                  # points [B, 4, N] on the device or in (pinned) host memory

The FPS start indices are still drawn on the CPU generator for every call, exactly like the reference
(pointnet_util.py:75), and reach the graph's static buffer through a small ring of pinned staging buffers.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch


class GraphedSemSeg:
    """Shape-keyed CUDA-graph cache around PointNet2SemSeg.forward."""

    RING = 8

    def __init__(self, net, warmup: int = 2):
        self.net = net.module if hasattr(net, "module") else net
        self.warmup = warmup
        self._graphs: Dict[Tuple, dict] = {}

    def _build(self, points: torch.Tensor, to_host: bool = False) -> dict:
        net, dev = self.net, points.device
        B, C, N = points.shape
        sizes = [N] + [m.npoint for m in (net.sa1, net.sa2, net.sa3)]
        st = {
            "x": torch.empty((B, C, N), dtype=torch.float32, device=dev),
            "starts": torch.zeros((len(sizes), B), dtype=torch.int64, device=dev),
            "sizes": sizes,
            "pinned": [torch.zeros((len(sizes), B), dtype=torch.int64).pin_memory() for _ in range(self.RING)],
            "events": [None] * self.RING,
            "slot": 0,
        }
        st["x"].copy_(points)
        classes = net.conv2.out_channels
        st["host_out"] = torch.empty((B, N, classes), dtype=torch.float32).pin_memory() if to_host else None
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):             # folds BatchNorm, packs weights, sizes the allocator pools
                net(st["x"], fps_starts=list(st["starts"].unbind(0)), host_out=st["host_out"])
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph), torch.no_grad():
            st["out"] = net(st["x"], fps_starts=list(st["starts"].unbind(0)), host_out=st["host_out"])
        st["graph"] = graph
        return st

    @torch.no_grad()
    def __call__(self, points: torch.Tensor, out: torch.Tensor = None, to_host: bool = False) -> torch.Tensor:
        """Replays the captured forward.  Returns the graph's static output buffer (valid until the next call)
        or, when `out` is given (device or pinned host), copies the log-probabilities there asynchronously.
        to_host=True: the device-to-host copies are nodes of the graph (the last level runs in two batch halves
        and the first half travels while the second is computed); returns the runner's static PINNED HOST buffer,
        complete once the current stream has been synchronised and valid until the next call."""
        dev = points.device if points.is_cuda else torch.device("cuda", torch.cuda.current_device())
        key = (tuple(points.shape), dev, bool(to_host))
        st = self._graphs.get(key)
        if st is None:
            st = self._graphs[key] = self._build(points.to(dev), to_host=bool(to_host))
        B = points.shape[0]
        slot = st["slot"] = (st["slot"] + 1) % self.RING
        if st["events"][slot] is not None:
            st["events"][slot].synchronize()          # the copy that last used this staging buffer has run
        pinned = st["pinned"][slot]
        for i, n in enumerate(st["sizes"]):           # the reference's draws, same generator, same order
            pinned[i] = torch.randint(0, n, (B,), dtype=torch.long)
        st["starts"].copy_(pinned, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        st["events"][slot] = ev
        st["x"].copy_(points, non_blocking=True)
        st["graph"].replay()
        if to_host:
            return st["host_out"]
        if out is not None:
            out.copy_(st["out"], non_blocking=True)
            return out
        return st["out"]
