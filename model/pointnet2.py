"""Drop-in alias: `model.pointnet2` as the reference spells it (see pointnet12_b200/model/pointnet2.py)."""
from pointnet12_b200.model.pointnet2 import *  # noqa: F401,F403
