"""Drop-in alias: `model.pointnet_util` as the reference spells it (see pointnet12_b200/model/pointnet_util.py)."""
from pointnet12_b200.model.pointnet_util import *  # noqa: F401,F403
