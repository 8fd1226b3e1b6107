"""Drop-in alias: `model.chamfer` as the reference spells it (see pointnet12_b200/model/chamfer.py)."""
from pointnet12_b200.model.chamfer import chamfer_batch, chamfer_non_batch  # noqa: F401
