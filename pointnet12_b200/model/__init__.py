"""Reference-shaped API: pointnet12_b200.model.{pointnet_util, pointnet2, pointnet, utils}."""
