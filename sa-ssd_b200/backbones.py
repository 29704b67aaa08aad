"""``mmdet.models.backbones`` mirror for the hot path: SimpleVoxel
(mmdet/models/backbones/vxnet.py:99-116)."""
import torch
from torch import nn

from . import ops


class SimpleVoxel(nn.Module):
    """Mean of the points of every voxel.  Same constructor kwargs as the reference
    (configs/car_cfg.py:3-8); ``use_norm``/``num_filters``/``with_distance`` are accepted
    and, as in the reference, unused."""

    def __init__(self, num_input_features=4, use_norm=True, num_filters=(32, 128), with_distance=False,
                 name="VoxelFeatureExtractor"):
        super().__init__()
        self.name = name
        self.num_input_features = num_input_features
        if num_input_features != 4:
            raise NotImplementedError("SA-SSD configs use 4 point features (x, y, z, r)")

    def forward(self, features, num_voxels, d_rows=None):
        """features [M, max_points, 4] f32, num_voxels [M] int -> [M, 4]."""
        ops.require_cuda()
        return ops.voxel_mean(features.contiguous(), num_voxels.to(torch.int32).contiguous(), d_rows)
