"""pointnet12_b200 -- B200-native PointNet / PointNet++ forward path behind the reference's Python API.

    from pointnet12_b200.model.pointnet2 import PointNet2SemSeg
    from pointnet12_b200.model.utils import load_pointnet

All computation runs in hand-written sm_100a CUDA kernels (pointnet12_b200/csrc) reached through the
C ABI declared in include/pn12_b200.h; there is no CPU fallback.
"""
__version__ = "0.1.0"
