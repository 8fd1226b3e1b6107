"""Drop-in alias: `model.pointnet` as the reference spells it (see pointnet12_b200/model/pointnet.py)."""
from pointnet12_b200.model.pointnet import *  # noqa: F401,F403
