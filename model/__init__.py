"""Reference-compatible import path (`from model.pointnet2 import PointNet2SemSeg`): aliases of pointnet12_b200.model."""
