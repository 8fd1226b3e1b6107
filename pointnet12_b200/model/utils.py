"""B200-native drop-in for the reference's model/utils.py: build a net, load a checkpoint, eval().

The reference wraps the model in torch.nn.DataParallel (model/utils.py:22), which is why its
checkpoints carry a `module.` prefix on every key.  Here scaling is one process per GPU, so the model
is returned inside a thin `ModuleWrapper` that owns it as `.module`: state_dict keys keep the
`module.` prefix and the reference's checkpoints load with strict=True, unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .pointnet import PointNetSeg
from .pointnet2 import PointNet2SemSeg


class ModuleWrapper(nn.Module):
    """DataParallel-shaped holder (`.module`, `module.`-prefixed state_dict) without the scatter/gather."""

    def __init__(self, module: nn.Module):
        super().__init__()
        self.module = module

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def load_pointnet(model_name, num_classes, fn_pth, device=None):
    """Reference model/utils.py:15-34.  `device` defaults to the current CUDA device (required: no CPU path)."""
    if model_name == 'pointnet':
        model = PointNetSeg(num_classes, input_dims=4, feature_transform=True)
    else:
        model = PointNet2SemSeg(num_classes, feature_dims=1)
    model = ModuleWrapper(model)
    assert fn_pth is not None, 'No pretrain model'
    if not torch.cuda.is_available():
        raise RuntimeError("load_pointnet: no CUDA device; pointnet12_b200 has no CPU fallback")
    device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    checkpoint = torch.load(fn_pth, map_location=device)
    model.load_state_dict(checkpoint)
    model.to(device)
    model.eval()
    return model
