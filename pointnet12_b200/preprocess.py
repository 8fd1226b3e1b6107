"""Raw SemanticKITTI scans -> [B, 4, N] network input and labels, on the device.

Replaces the host path of the reference between the dataset files and the model: Semantic_KITTI_Utils.get
(data_utils/kitti_utils.py:183-227: learning-map lookup, drop class 0, label - 1, in-view field-of-view filter) and
SemKITTI_Loader.__getitem__ (data_utils/SemKITTI_Loader.py:91-115: pcd_normalize, pcd_jitter when training,
np.random.choice(length, npoints, replace=True)) plus the DataLoader collate and pcdseg.py:167's transpose.  The wire
format in is the dataset's own (.bin float32 x 4, .label uint32); out come the tensors the model and the loss take.

    pre = ScanPreprocessor(learning_map)                         # dict {raw id: class 0..19} from config/semantic-kitti.yaml
    batch = pre.upload(list_of_points_M4, list_of_raw_labels)    # pinned staging + one H2D copy per array
    points, labels = pre(batch, npoints=8000, train=True)        # [B,4,N] float32, [B,N] int64 on the device

Randomness: drawn on the device from a Philox stream seeded from torch's generator state (the reference uses numpy's global
generator, which a GPU cannot replay); `choice=` / `noise=` inject the draws instead -- the parity tests pass the
reference's own draws and get its output bit for bit.
"""
from __future__ import annotations

from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nv
from . import ops


class ScanBatch:
    """B raw scans concatenated on the device."""

    __slots__ = ("points", "raw_label", "offsets", "lengths", "B", "max_points")


class ScanPreprocessor:
    H_FOV = (-40.0, 40.0)        # Semantic_KITTI_Utils.get -> set_filter([-40, 40], [-20, 20]) for subset 'inview'
    V_FOV = (-20.0, 20.0)
    SIGMA, CLIP = 0.01, 0.05     # pcd_jitter defaults

    def __init__(self, learning_map: Dict[int, int], subset: str = "inview", device=None):
        if subset not in ("inview", "all"):
            raise ValueError("subset must be 'inview' or 'all'")
        self.subset = subset
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        lut = np.zeros(max(learning_map) + 1, dtype=np.uint8)
        for k, v in learning_map.items():
            lut[int(k)] = int(v)
        self.lut_host = lut
        self.lut = torch.from_numpy(lut).to(self.device)
        # the comparisons of hv_in_range run in float32 (numpy casts the python scalar to the array's dtype)
        self.bounds = (float(np.float32(-self.H_FOV[1] * np.pi / 180)), float(np.float32(-self.H_FOV[0] * np.pi / 180)),
                       float(np.float32(self.V_FOV[0] * np.pi / 180)), float(np.float32(self.V_FOV[1] * np.pi / 180)))
        self._calls = 0
        self._stage = {}                 # grow-only pinned staging buffers (pinning memory costs milliseconds per call otherwise)
        self._seed_ring, self._seed_slot = None, 0

    def upload(self, points: Sequence[np.ndarray], raw_labels: Sequence[np.ndarray]) -> ScanBatch:
        if len(points) != len(raw_labels) or not points:
            raise ValueError("need one label array per scan")
        lengths = [int(p.shape[0]) for p in points]
        for p, l in zip(points, raw_labels):
            if p.ndim != 2 or p.shape[1] != 4 or l.shape[0] != p.shape[0]:
                raise ValueError("Scan and Label don't contain same number of points")     # kitti_utils.py:210
        total = sum(lengths)
        # pinned staging, reused across calls once the copy that last read it has completed
        st = self._stage
        if st.get("cap", 0) < total:
            cap = int(total * 1.25) + 1024
            st.update(cap=cap, points=torch.empty((cap, 4), dtype=torch.float32).pin_memory(),
                      labels=torch.empty((cap,), dtype=torch.int32).pin_memory(), event=None)
        if st["event"] is not None:
            st["event"].synchronize()
        hp, hl = st["points"][:total], st["labels"][:total]
        o = 0
        for p, l, n in zip(points, raw_labels, lengths):
            hp[o:o + n] = torch.from_numpy(np.ascontiguousarray(p, dtype=np.float32))
            hl[o:o + n] = torch.from_numpy(np.ascontiguousarray(l).astype(np.uint32).view(np.int32))
            o += n
        b = ScanBatch()
        b.points = hp.to(self.device, non_blocking=True)
        b.raw_label = hl.to(self.device, non_blocking=True)
        b.offsets = torch.tensor([0] + list(np.cumsum(lengths)), dtype=torch.int64).to(self.device)
        st["event"] = torch.cuda.Event()
        st["event"].record(torch.cuda.current_stream(self.device))
        b.lengths, b.B, b.max_points = lengths, len(lengths), max(lengths)
        return b

    def _seed(self) -> torch.Tensor:
        """{seed, offset} of this call's Philox stream on the device, sent through a small ring of pinned buffers (no
        synchronous pageable copy on the way)."""
        if self._seed_ring is None:
            self._seed_ring = [(torch.zeros((2,), dtype=torch.int64).pin_memory(), torch.zeros((2,), dtype=torch.int64, device=self.device),
                                [None]) for _ in range(8)]
        self._seed_slot = (self._seed_slot + 1) % len(self._seed_ring)
        host, dev, ev = self._seed_ring[self._seed_slot]
        if ev[0] is not None:
            ev[0].synchronize()
        self._calls += 1
        host[0] = torch.initial_seed() & 0x7FFFFFFFFFFFFFFF
        host[1] = self._calls
        dev.copy_(host, non_blocking=True)
        ev[0] = torch.cuda.Event()
        ev[0].record(torch.cuda.current_stream(self.device))
        return dev

    def filter(self, batch: ScanBatch) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (kept int32 [total]: per scan the indices of its kept points in file order, kept_count int32 [B])."""
        total = batch.points.shape[0]
        kept = torch.empty((total,), dtype=torch.int32, device=self.device)
        count = torch.empty((batch.B,), dtype=torch.int32, device=self.device)
        nbytes = int(nv.lib().pn_scan_workspace_bytes(batch.B, batch.max_points))
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=self.device)
        with ops._on_device(kept):
            nv.call("pn_scan_filter_f32", batch.points.data_ptr(), batch.raw_label.data_ptr(), batch.offsets.data_ptr(), batch.B,
                    batch.max_points, self.lut.data_ptr(), self.lut.numel(), int(self.subset == "inview"), *self.bounds,
                    kept.data_ptr(), count.data_ptr(), ws.data_ptr(), nbytes, ops._stream())
        return kept, count

    def __call__(self, batch: ScanBatch, npoints: int, train: bool, choice: Optional[torch.Tensor] = None,
                 noise: Optional[torch.Tensor] = None, filtered=None, check_empty: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
        """-> (points [B, 4, npoints] float32, labels [B, npoints] int64).
        choice int64 [B, npoints] / noise float32 [total, 4] (row offsets[b] + i = jitter of the i-th kept point of scan
        b): the reference's own draws, for parity; by default both are drawn on the device (jitter only when train).
        check_empty: a scan whose filter keeps no point raises ValueError, like the reference's np.random.choice(0, npoints)
        (SemKITTI_Loader.py:100-106); the check reads the B kept counts back (one small synchronising copy) -- a throughput
        loop that knows its scans may pass False, the sampling kernel then emits all-zero points with label 0 for such a scan."""
        kept, count = self.filter(batch) if filtered is None else filtered
        if check_empty:
            empty = (count == 0).nonzero().flatten().tolist()
            if empty:
                raise ValueError(f"scan(s) {empty} of the batch keep no point after the '{self.subset}' filter: nothing to sample from")
        out = torch.empty((batch.B, 4, int(npoints)), dtype=torch.float32, device=self.device)
        labels = torch.empty((batch.B, int(npoints)), dtype=torch.int64, device=self.device)
        seed = None
        if choice is None or (train and noise is None):
            seed = self._seed()
        if choice is not None:
            choice = ops._i64(choice, "choice")
        if noise is not None:
            noise = ops._f32(noise, "noise").contiguous()
        sigma = self.SIGMA if (train and noise is None) else 0.0
        with ops._on_device(out):
            nv.call("pn_scan_sample_f32", batch.points.data_ptr(), batch.raw_label.data_ptr(), batch.offsets.data_ptr(), batch.B,
                    self.lut.data_ptr(), self.lut.numel(), kept.data_ptr(), count.data_ptr(), int(npoints), ops._p(choice),
                    ops._p(noise), float(sigma), float(self.CLIP), ops._p(seed), out.data_ptr(), labels.data_ptr(), ops._stream())
        return out, labels


class SemKITTI_2_Common:
    """Drop-in for data_utils/kitti_utils.py:61-117: wraps a model and merges its 19 SemanticKITTI log-probability channels into
    the 16 'common' classes (max over the merged pair) with one kernel instead of 16 indexed copies.  Same constructor,
    attributes (`semkitti_names`, `semkitti_2_common`, `colors`, `semkitti_colors`, `common`) and call as the reference."""

    semkitti_names = ['car', 'bicycle', 'motorcycle', 'truck', 'other-vehicle', 'person', 'bicyclist', 'motorcyclist', 'road',
                      'parking', 'sidewalk', 'other-ground', 'building', 'fence', 'vegetation', 'trunk', 'terrain', 'pole',
                      'traffic-sign']
    semkitti_2_common = ['road', 'parking+sidewalk', 'building', 'fence', 'trunk+pole', 'traffic-sign', 'vegetation', 'terrain',
                         'person', 'bicyclist+motorcyclist', 'car', 'truck', 'other-vehicle', 'motorcycle', 'bicycle', 'other-ground']
    _colors = [[245, 150, 100], [245, 230, 100], [150, 60, 30], [180, 30, 80], [255, 0, 0], [30, 30, 255], [200, 40, 255],
               [90, 30, 150], [255, 0, 255], [255, 150, 255], [75, 0, 75], [75, 0, 175], [0, 200, 255], [50, 120, 255],
               [0, 175, 0], [0, 60, 135], [80, 240, 150], [150, 240, 255], [0, 0, 255]]

    def __init__(self, model, model_name):
        self.model_name, self.model, self.common = model_name, model, None
        pairs = [[self.semkitti_names.index(n) for n in c.split('+')] for c in self.semkitti_2_common]
        if any(len(p) > 2 for p in pairs):
            raise NotImplementedError("not implemented!")
        self._src = [[p[0] for p in pairs], [p[-1] for p in pairs]]
        self.semkitti_colors = np.array(self._colors)
        self.colors = np.array([self._colors[p[0]] for p in pairs])
        self._dev_src = {}

    def merge(self, logits: torch.Tensor) -> torch.Tensor:
        """[B, N, 19] -> [B, N, 16]."""
        ops._need_cuda(logits, "logits")
        x = logits.float().contiguous()
        k = x.shape[-1]
        if x.device not in self._dev_src:
            self._dev_src[x.device] = torch.tensor(self._src, dtype=torch.int32, device=x.device)
        src = self._dev_src[x.device]
        out = torch.empty(tuple(x.shape[:-1]) + (len(self.semkitti_2_common),), dtype=torch.float32, device=x.device)
        rows = x.numel() // k
        with ops._on_device(x):
            nv.call("pn_class_merge_f32", x.data_ptr(), k, rows, k, out.shape[-1], src[0].data_ptr(), src[1].data_ptr(),
                    out.data_ptr(), ops._stream())
        return out

    def __call__(self, x):
        if self.model_name == 'pointnet':
            logits, feature_transform = self.model(x)
        else:
            logits = self.model(x)
        self.common = self.merge(logits)
        return (self.common, feature_transform) if self.model_name == 'pointnet' else self.common
