"""Drop-in alias: `model.utils` as the reference spells it (see pointnet12_b200/model/utils.py)."""
from pointnet12_b200.model.utils import *  # noqa: F401,F403
