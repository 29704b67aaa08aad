"""Anchor grid and anchors_mask: mmdet/core/anchor/anchor3d_generator.py:3-41,81-103,
mmdet/core/bbox3d/geometry.py:404-426,675-709, mmdet/datasets/kitti.py:80-88,333-343.

The anchor grid and each anchor's integer footprint are static (host, numpy fp32 — the
same arithmetic as the reference, computed once); the per-frame mask is a CUDA kernel
(csrc/voxelize.cu: count map -> 2-D prefix sum -> 4-corner lookup)."""
import numpy as np
import torch

from . import ops


def create_anchors_3d_stride(feature_size, sizes=(1.6, 3.9, 1.56), anchor_strides=(0.4, 0.4, 0.0),
                             anchor_offsets=(0.2, -39.8, -1.78), rotations=(0, np.pi / 2), dtype=np.float32):
    """Returns [D, H, W, num_sizes, num_rots, 7] anchors (x, y, z, w, l, h, r), fp32 centres
    ``arange(n) * stride + offset`` exactly as the reference computes them."""
    D, H, W = [int(v) for v in feature_size]
    xs, ys, zs = [dtype(v) for v in anchor_strides]
    xo, yo, zo = [dtype(v) for v in anchor_offsets]
    zc = np.arange(D, dtype=dtype) * zs + zo
    yc = np.arange(H, dtype=dtype) * ys + yo
    xc = np.arange(W, dtype=dtype) * xs + xo
    sizes = np.asarray(sizes, dtype=dtype).reshape(-1, 3)
    rots = np.asarray(rotations, dtype=dtype)
    out = np.empty((D, H, W, sizes.shape[0], rots.shape[0], 7), dtype=dtype)
    out[..., 0] = xc[None, None, :, None, None]
    out[..., 1] = yc[None, :, None, None, None]
    out[..., 2] = zc[:, None, None, None, None]
    out[..., 3:6] = sizes[None, None, None, :, None, :]
    out[..., 6] = rots[None, None, None, None, :]
    return out


class AnchorGeneratorStride:
    """Constructor kwargs as in configs/car_cfg.py:123-130."""

    def __init__(self, sizes=(1.6, 3.9, 1.56), anchor_strides=(0.4, 0.4, 1.0), anchor_offsets=(0.2, -39.8, -1.78),
                 rotations=(0, np.pi / 2), dtype=np.float32):
        self._sizes, self._anchor_strides = sizes, anchor_strides
        self._anchor_offsets, self._rotations, self._dtype = anchor_offsets, rotations, dtype

    @property
    def num_anchors_per_localization(self):
        return len(self._rotations) * np.asarray(self._sizes).reshape(-1, 3).shape[0]

    def __call__(self, feature_map_size):
        return create_anchors_3d_stride(feature_map_size, self._sizes, self._anchor_strides, self._anchor_offsets,
                                        self._rotations, self._dtype)


def limit_period(val, offset=0.5, period=np.pi):
    return val - np.floor(val / period + offset) * period


def rbbox2d_to_near_bbox(rbboxes):
    """[N,5] (x, y, xdim, ydim, rad) -> nearest axis-aligned [N,4] (geometry.py:414-426)."""
    rots = rbboxes[..., -1]
    swap = (np.abs(limit_period(rots, 0.5, np.pi)) > np.pi / 4)[..., None]
    ctr = np.where(swap, rbboxes[:, [0, 1, 3, 2]], rbboxes[:, :4])
    return np.concatenate([ctr[:, :2] - ctr[:, 2:] / 2, ctr[:, :2] + ctr[:, 2:] / 2], axis=-1)


def anchor_rects(anchors_bv, voxel_size, pc_range, grid_size):
    """Integer footprint (c0, c1, c2, c3) of each anchor on the voxel grid, clamped as in
    geometry.py:691-702 (fp32 subtract / divide / floor)."""
    vs = np.asarray(voxel_size, np.float32)
    off = np.asarray(pc_range, np.float32)
    a = np.asarray(anchors_bv, np.float32)
    W, H = int(grid_size[0]), int(grid_size[1])
    c0 = np.maximum(np.floor((a[:, 0] - off[0]) / vs[0]).astype(np.int32), 0)
    c1 = np.maximum(np.floor((a[:, 1] - off[1]) / vs[1]).astype(np.int32), 0)
    c2 = np.minimum(np.floor((a[:, 2] - off[0]) / vs[0]).astype(np.int32), W - 1)
    c3 = np.minimum(np.floor((a[:, 3] - off[1]) / vs[1]).astype(np.int32), H - 1)
    return np.ascontiguousarray(np.stack([c0, c1, c2, c3], 1).astype(np.int32))


class AnchorSet:
    """Static per-config anchor data: anchors [Na,7] (classes concatenated, kitti.py:85-88),
    their BEV boxes and integer footprints, resident on the device."""

    def __init__(self, anchor_generators, voxel_generator, out_size_factor=8, anchor_area_threshold=1, device=None):
        grid = voxel_generator.grid_size
        fms = [*(grid[:2] // out_size_factor), 1][::-1]                      # kitti.py:82-83
        gens = list(anchor_generators.values()) if isinstance(anchor_generators, dict) else list(anchor_generators)
        self.feature_map_size = [int(v) for v in fms]
        self.anchors = np.concatenate([g(fms).reshape(-1, 7) for g in gens], 0).astype(np.float32)
        self.anchors_bv = rbbox2d_to_near_bbox(self.anchors[..., [0, 1, 3, 4, 6]])
        self.rects = anchor_rects(self.anchors_bv, voxel_generator.voxel_size, voxel_generator.point_cloud_range, grid)
        self.grid_hw = (int(grid[1]), int(grid[0]))
        self.threshold = anchor_area_threshold
        self.device = device
        self._dev = None

    def to(self, device):
        self.device = device
        self._dev = None
        return self

    def device_tensors(self):
        if self._dev is None:
            self._dev = (torch.from_numpy(self.anchors).to(self.device), torch.from_numpy(self.rects).to(self.device))
        return self._dev

    def mask_device(self, coors, d_rows, batch):
        """coors [cap,4] (b,z,y,x) device, d_rows [1] -> mask [batch, Na] uint8 (no sync)."""
        _, rects = self.device_tensors()
        H, W = self.grid_hw
        return ops.anchor_mask(coors, d_rows, batch, H, W, rects, self.threshold)

    def mask(self, coordinates_zyx):
        """Per-frame numpy API (kitti.py:333-343): coordinates [M,3] (z,y,x) -> bool [Na]."""
        ops.require_cuda()
        c = torch.from_numpy(np.ascontiguousarray(coordinates_zyx, np.int32)).to(self.device)
        c4 = torch.nn.functional.pad(c, (1, 0), value=0).contiguous()
        d_rows = torch.tensor([c4.shape[0]], dtype=torch.int32, device=self.device)
        if c4.shape[0] == 0:
            c4 = torch.zeros((1, 4), dtype=torch.int32, device=self.device)
        return self.mask_device(c4, d_rows, 1)[0].bool().cpu().numpy()
