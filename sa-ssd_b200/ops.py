"""Functional wrappers over the C ABI (include/sassd_b200.h) on torch CUDA tensors.

torch is plumbing here (device memory, streams); every computation is a
hand-written sm_100a kernel in csrc/.  All wrappers are asynchronous on the
current torch stream and keep data-dependent sizes on the device (``d_rows``
style int32 tensors) — nothing in this module synchronises.
"""
import ctypes
import os
import math

import numpy as np
import torch

from . import lib as _lib
from .lib import Conv2dDesc, SpconvDesc  # noqa: E402
from .lib import (GCONV_CONV2D, GCONV_ROWS, GCONV_TABLE, PREC_F16X3, PREC_FP32, PREC_TF32X3, GConvDesc, VoxelParams,
                  check)

NMS_CAP = 4096
# Product default: every conv on the tcgen05 kernels with the fp32-accurate 3xFP16 operand split (csrc/spconv_split.cu,
# csrc/conv2d_tma.cu).  PREC_FP32 (CUDA-core FFMA) and PREC_TF32X3 stay selectable per model
# (SingleStageDetector.set_precision) for bisecting.
DEFAULT_PRECISION = PREC_F16X3


def _L():
    return _lib.load()


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    assert t.is_cuda and t.is_contiguous(), "sassd ops need contiguous CUDA tensors"
    return ctypes.c_void_p(t.data_ptr())


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.SassdError("sassd_b200 needs a CUDA device (sm_100a); there is no CPU fallback")


def next_pow2(n):
    return 1 << max(1, int(math.ceil(math.log2(max(2, n)))))


# --- launch accounting / optional per-call CUDA-event timing (bench.py) ---------------
# kernels launched by each C-ABI entry point (memsets not counted)
_KERNELS = {"sassd_voxelize": 4, "sassd_voxel_mean": 1, "sassd_anchor_mask": 4, "sassd_hash_build": 1,
            "sassd_rulebook_subm": 1, "sassd_rulebook_conv_outputs": 2, "sassd_rulebook_conv_outputs_hash": 2, "sassd_rulebook_conv_nbr": 1,
            "sassd_rulebook_pairs": 1, "sassd_gconv": 1, "sassd_gconv_pack": 1, "sassd_spconv_pack": 1, "sassd_rotate_overlap_eval": 1, "sassd_conv2d_f16x3": 1, "sassd_conv2d_f16x3_occ": 1, "sassd_spconv_f16x3": 1, "sassd_features_to_split": 1, "sassd_split_rows_to_bev": 1, "sassd_sparse_to_bev_split": 1, "sassd_sparse_to_bev": 1, "sassd_decode_select": 2,
            "sassd_pswarp": 1, "sassd_rescore_nms": 3, "sassd_nms_mask": 1, "sassd_nms_sorted": 2,
            "sassd_boxes_iou_bev": 1}
LAUNCHES = 0          # running count of kernels launched through this module
PROFILE = None        # set to a list to collect (name, label, start_event, end_event)


def _call(name, label, *args):
    global LAUNCHES
    LAUNCHES += _KERNELS[name]
    fn = getattr(_L(), name)
    if PROFILE is None:
        check(fn(*args), name)
        return
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    check(fn(*args), name)
    e1.record()
    PROFILE.append((name, label, e0, e1))


class Workspace:
    """Grow-only byte buffers keyed by purpose (the C ABI never allocates)."""

    def __init__(self):
        self._bufs = {}

    def get(self, key, nbytes, device):
        buf = self._bufs.get(key)
        if buf is None or buf.numel() < nbytes or buf.device != device:
            buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self._bufs[key] = buf
        return buf


_WS = Workspace()


# ---------------------------------------------------------------------------- voxelize
def make_voxel_params(voxel_size, pc_range, max_points, max_voxels):
    vs = np.asarray(voxel_size, np.float32)
    rg = np.asarray(pc_range, np.float32)
    grid = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)   # voxel_generator.py:13-15
    p = VoxelParams()
    for j in range(3):
        p.voxel_size[j] = float(vs[j]); p.range_min[j] = float(rg[j]); p.grid[j] = int(grid[j])
    p.max_points = int(max_points); p.max_voxels = int(max_voxels)
    return p, grid


def voxelize(points, pt_off, batch, params, rows_cap, slots_per_frame, status, want_mean=True, ws=None):
    """points [Ncap,4] f32, pt_off [batch+1] i32 (device).  Returns capacity-sized
    (voxels, coors, num_points, mean, frame_rows[batch+1])."""
    dev = points.device
    n_cap = points.shape[0]
    voxels = torch.empty((rows_cap, params.max_points, 4), dtype=torch.float32, device=dev)
    coors = torch.empty((rows_cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((rows_cap,), dtype=torch.int32, device=dev)
    mean = torch.empty((rows_cap, 4), dtype=torch.float32, device=dev) if want_mean else None
    frame_rows = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    nbytes = _L().sassd_voxelize_workspace_bytes(n_cap, batch, slots_per_frame)
    w = (ws or _WS).get("voxelize", nbytes, dev)
    _call("sassd_voxelize", None, _ptr(points), _ptr(pt_off), n_cap, batch, ctypes.byref(params), slots_per_frame,
                              _ptr(voxels), _ptr(coors), _ptr(num), _ptr(mean), rows_cap, _ptr(frame_rows),
                              _ptr(status), _ptr(w), w.numel(), _stream())
    return voxels, coors, num, mean, frame_rows


def voxel_mean(voxels, num_points, d_rows=None):
    rows, maxp = voxels.shape[0], voxels.shape[1]
    mean = torch.empty((rows, 4), dtype=torch.float32, device=voxels.device)
    _call("sassd_voxel_mean", None, _ptr(voxels), _ptr(num_points), _ptr(d_rows), rows, maxp, _ptr(mean), _stream())
    return mean


def anchor_mask(coors, d_rows, batch, H, W, rects, threshold=1, ws=None):
    dev = coors.device
    na = rects.shape[0]
    mask = torch.empty((batch, na), dtype=torch.uint8, device=dev)
    nbytes = _L().sassd_anchor_mask_workspace_bytes(batch, H, W)
    w = (ws or _WS).get("amask", nbytes, dev)
    _call("sassd_anchor_mask", None, _ptr(coors), _ptr(d_rows), coors.shape[0], batch, H, W, _ptr(rects), na,
                                 int(threshold), _ptr(mask), _ptr(w), w.numel(), _stream())
    return mask


# ---------------------------------------------------------------------------- rulebooks
class HashIndex:
    def __init__(self, rows_cap, device):
        self.slots = next_pow2(2 * max(rows_cap, 1))
        self.keys = torch.empty((self.slots,), dtype=torch.int32, device=device)
        self.vals = torch.empty((self.slots,), dtype=torch.int32, device=device)


def hash_build(index, coors, d_rows, batch, shape, status):
    D, H, W = shape
    _call("sassd_hash_build", None, _ptr(coors), _ptr(d_rows), coors.shape[0], batch, D, H, W, _ptr(index.keys),
                                _ptr(index.vals), index.slots, _ptr(status), _stream())
    return index


def _tile_mask_buffer(rows_cap, device):
    """int32 per 128-row tile: which of the 27 taps occur in the tile (written by the rulebook kernels, read by
    spconv_split to skip absent taps)."""
    n = (rows_cap + _lib.SPCONV_TILE_ROWS - 1) // _lib.SPCONV_TILE_ROWS
    return torch.empty((max(n, 1),), dtype=torch.int32, device=device)


def rulebook_subm(coors, d_rows, shape, index, nbr=None):
    """Returns (nbr [rows_cap, 27], tile_mask [tiles])."""
    D, H, W = shape
    rows_cap = coors.shape[0]
    if nbr is None:
        nbr = torch.empty((rows_cap, 27), dtype=torch.int32, device=coors.device)
    tmask = _tile_mask_buffer(rows_cap, coors.device)
    _call("sassd_rulebook_subm", None, _ptr(coors), _ptr(d_rows), rows_cap, D, H, W, _ptr(index.keys), _ptr(index.vals),
                                   index.slots, _ptr(nbr), _ptr(tmask), _stream())
    return nbr, tmask


def conv_out_shape(shape):
    return [(s + 2 - 3) // 2 + 1 for s in shape]


def rulebook_conv_outputs(coors_in, d_rows_in, batch, shape, rows_cap_out, status, ws=None, ws_key="rbconv",
                          index_out=None):
    """Active output set of a strided (k3,s2,p1) conv, sorted by flattened index.  With ``index_out`` (a fresh
    HashIndex of the output level) the rows are hashed as they are emitted.  Returns coors_out [cap,4], d_rows_out [1],
    out_shape."""
    dev = coors_in.device
    D, H, W = shape
    Do, Ho, Wo = conv_out_shape(shape)
    coors_out = torch.empty((rows_cap_out, 4), dtype=torch.int32, device=dev)
    d_rows_out = torch.empty((1,), dtype=torch.int32, device=dev)
    nbytes = _L().sassd_rulebook_conv_workspace_bytes(batch, Do, Ho, Wo)
    w = (ws or _WS).get(ws_key, nbytes, dev)
    if index_out is None:
        _call("sassd_rulebook_conv_outputs", None, _ptr(coors_in), _ptr(d_rows_in), coors_in.shape[0], batch, D, H, W,
              _ptr(coors_out), _ptr(d_rows_out), rows_cap_out, _ptr(status), _ptr(w), w.numel(), _stream())
    else:
        _call("sassd_rulebook_conv_outputs_hash", None, _ptr(coors_in), _ptr(d_rows_in), coors_in.shape[0], batch, D, H, W,
              _ptr(coors_out), _ptr(d_rows_out), rows_cap_out, _ptr(index_out.keys), _ptr(index_out.vals),
              index_out.slots, _ptr(status), _ptr(w), w.numel(), _stream())
    return coors_out, d_rows_out, [Do, Ho, Wo]


def rulebook_conv_nbr(coors_out, d_rows_out, shape_in, index_in):
    """Neighbour table of the strided conv (probes the INPUT level's hash).  Returns nbr [cap,27], tile_mask."""
    D, H, W = shape_in
    rows_cap_out = coors_out.shape[0]
    nbr = torch.empty((rows_cap_out, 27), dtype=torch.int32, device=coors_out.device)
    tmask = _tile_mask_buffer(rows_cap_out, coors_out.device)
    _call("sassd_rulebook_conv_nbr", None, _ptr(coors_out), _ptr(d_rows_out), rows_cap_out, D, H, W, _ptr(index_in.keys),
          _ptr(index_in.vals), index_in.slots, _ptr(nbr), _ptr(tmask), _stream())
    return nbr, tmask


def rulebook_conv(coors_in, d_rows_in, batch, shape, index_in, rows_cap_out, status, ws=None, ws_key="rbconv",
                  index_out=None):
    """Strided (k3,s2,p1) rulebook.  Returns coors_out [cap,4], d_rows_out [1], nbr [cap,27], out_shape, tile_mask."""
    coors_out, d_rows_out, so = rulebook_conv_outputs(coors_in, d_rows_in, batch, shape, rows_cap_out, status, ws, ws_key,
                                                      index_out)
    nbr, tmask = rulebook_conv_nbr(coors_out, d_rows_out, shape, index_in)
    return coors_out, d_rows_out, nbr, so, tmask


def rulebook_pairs(nbr, d_rows):
    rows_cap = nbr.shape[0]
    pairs = torch.empty((2, 27, rows_cap), dtype=torch.int32, device=nbr.device)
    num = torch.empty((27,), dtype=torch.int32, device=nbr.device)
    _call("sassd_rulebook_pairs", None, _ptr(nbr), _ptr(d_rows), rows_cap, _ptr(pairs), _ptr(num), _stream())
    return pairs, num


# ---------------------------------------------------------------------------- gathered conv
_TC_PACKS = {}


def pack_tc(weight, precision):
    """weight [taps, cin, cout] fp32 (device) -> tensor-core pack (hi/lo split, 128B-swizzled K-major blocks)."""
    taps, cin, cout = weight.shape
    nbytes = _L().sassd_gconv_pack_bytes(taps, cin, cout, precision)
    packed = torch.empty((nbytes,), dtype=torch.uint8, device=weight.device)
    _call("sassd_gconv_pack", None, _ptr(weight), taps, cin, cout, precision, _ptr(packed), _stream())
    return packed


def tc_pack_cached(weight, precision):
    """Pack once per weight tensor.  The entry keeps the source tensor alive so that its address cannot be
    recycled for a different weight while the pack is cached (bounded FIFO)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), precision)
    ent = _TC_PACKS.get(key)
    if ent is None:
        if len(_TC_PACKS) >= 256:
            _TC_PACKS.pop(next(iter(_TC_PACKS)))
        ent = (pack_tc(weight.contiguous(), precision), weight)
        _TC_PACKS[key] = ent
    return ent[0]


def spconv_pack_cached(weight, cin_stored):
    """Tap-packed fp16 hi/lo weight blocks for sassd_spconv_f16x3 (cached like tc_pack_cached)."""
    key = (weight.data_ptr(), weight._version, tuple(weight.shape), "spconv", cin_stored)
    ent = _TC_PACKS.get(key)
    if ent is None:
        if len(_TC_PACKS) >= 256:
            _TC_PACKS.pop(next(iter(_TC_PACKS)))
        w = weight.contiguous()
        taps, cin, cout = w.shape
        nbytes = _L().sassd_spconv_pack_bytes(taps, cin_stored, cout)
        if nbytes == 0:
            raise _lib.SassdError("sassd_spconv_pack_bytes: unsupported shape taps=%d cin_stored=%d cout=%d"
                             % (taps, cin_stored, cout))
        packed = torch.empty((nbytes,), dtype=torch.uint8, device=w.device)
        _call("sassd_spconv_pack", None, _ptr(w), taps, cin, cin_stored, cout, _ptr(packed), _stream())
        ent = (packed, weight)
        _TC_PACKS[key] = ent
    return ent[0]


def gconv(inp, weight, scale, shift, out, *, mode, taps, cin, cout, relu, nbr=None, d_rows=None, rows_cap=None,
          batch=0, H=0, W=0, precision=PREC_FP32):
    """out[m,:] = act((sum_t in[row(m,t),:] @ W[t]) * scale + shift); see sassd_b200.h."""
    if precision in (PREC_TF32X3, PREC_F16X3):
        weight = tc_pack_cached(weight, precision)
    d = GConvDesc()
    d.mode, d.precision = mode, precision
    d.cin, d.cout, d.taps = cin, cout, taps
    d.in_stride = inp.stride(-2) if inp.dim() >= 2 else cin
    d.out_stride = out.stride(-2) if out.dim() >= 2 else cout
    d.rows_cap = int(rows_cap if rows_cap is not None else out.numel() // d.out_stride)
    d.batch, d.H, d.W = batch, H, W
    d.relu = 1 if relu else 0
    label = "gconv[%s taps=%d %d->%d]" % (("table", "conv2d", "rows")[mode], taps, cin, cout)
    _call("sassd_gconv", label, ctypes.byref(d), _ptr_any(inp), _ptr(weight), _ptr(scale), _ptr(shift), _ptr(nbr),
                           _ptr(d_rows), _ptr_any(out), _stream())
    return out


def _ptr_any(t):
    assert t.is_cuda
    return ctypes.c_void_p(t.data_ptr())


def sparse_to_bev(feat, coors, d_rows, C, D, H, W, bev):
    _call("sassd_sparse_to_bev", None, _ptr(feat), _ptr(coors), _ptr(d_rows), feat.shape[0], C, D, H, W, _ptr(bev),
                                   _stream())
    return bev


# ---------------------------------------------------------------------------- head tail
def decode_select(head, num_class, anchors, mask, thr, k_cap, status, ws=None):
    """head [B,H,W,stride] NHWC; anchors [Na,7] (shared) or [B,Na,7] (per frame).
    Returns boxes [B,k_cap,7], labels, index, d_k [B]."""
    dev = head.device
    B, H, W, stride = head.shape
    per_frame = 1 if anchors.dim() == 3 else 0
    assert not per_frame or anchors.shape[0] == B
    na = anchors.shape[-2]
    boxes = torch.empty((B, k_cap, 7), dtype=torch.float32, device=dev)
    labels = torch.empty((B, k_cap), dtype=torch.int32, device=dev)
    index = torch.empty((B, k_cap), dtype=torch.int32, device=dev)
    d_k = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = _L().sassd_decode_select_workspace_bytes(B, na)
    w = (ws or _WS).get("decode", nbytes, dev)
    _call("sassd_decode_select", None, _ptr(head), stride, B, H, W, num_class, _ptr(anchors), per_frame, _ptr(mask), na,
                                   ctypes.c_float(thr), _ptr(boxes), _ptr(labels), _ptr(index), _ptr(d_k), k_cap,
                                   _ptr(status), _ptr(w), w.numel(), _stream())
    return boxes, labels, index, d_k


def pswarp(feat, boxes, d_k, off_x, off_y, spatial_scale):
    B, H, W, stride = feat.shape
    k_cap = boxes.shape[1]
    scores = torch.empty((B, k_cap), dtype=torch.float32, device=feat.device)
    _call("sassd_pswarp", None, _ptr(feat), stride, B, H, W, _ptr(boxes), _ptr(d_k), k_cap, ctypes.c_float(off_x),
                            ctypes.c_float(off_y), ctypes.c_float(spatial_scale), _ptr(scores), _stream())
    return scores


def rescore_nms(boxes, scores, labels, d_k, score_thr, iou_thr, det_cap, status, ws=None):
    dev = boxes.device
    B, k_cap = boxes.shape[0], boxes.shape[1]
    det = torch.empty((B, det_cap, 9), dtype=torch.float32, device=dev)
    d_ndet = torch.empty((B,), dtype=torch.int32, device=dev)
    nbytes = _L().sassd_rescore_nms_workspace_bytes(B, k_cap, NMS_CAP)
    w = (ws or _WS).get("nms", nbytes, dev)
    _call("sassd_rescore_nms", None, _ptr(boxes), _ptr(scores), _ptr(labels), _ptr(d_k), B, k_cap,
                                 ctypes.c_float(score_thr), ctypes.c_float(iou_thr), NMS_CAP, _ptr(det), _ptr(d_ndet),
                                 det_cap, _ptr(status), _ptr(w), w.numel(), _stream())
    return det, d_ndet


def nms_mask(boxes5, thr):
    n = boxes5.shape[0]
    colb = (n + 63) // 64
    mask = torch.zeros((n, max(colb, 1)), dtype=torch.int64, device=boxes5.device)
    _call("sassd_nms_mask", None, _ptr(boxes5), n, ctypes.c_float(thr), _ptr(mask), _stream())
    return mask[:, :colb]


def nms_sorted(boxes5, thr):
    """boxes sorted by score.  Returns (keep [n] int64 capacity-sized, d_nkeep [1])."""
    dev = boxes5.device
    n = boxes5.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    d_n = torch.zeros((1,), dtype=torch.int32, device=dev)
    nbytes = _L().sassd_nms_workspace_bytes(n)
    w = _WS.get("nms_sorted", nbytes, dev)
    _call("sassd_nms_sorted", None, _ptr(boxes5), n, ctypes.c_float(thr), _ptr(keep), _ptr(d_n), _ptr(w), w.numel(),
                                _stream())
    return keep, d_n


def boxes_iou_bev(a, b):
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _call("sassd_boxes_iou_bev", None, _ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream())
    return out


# ---------------------------------------------------------------------------- TMA dense conv on split maps
class SplitMap:
    """Activation map as two fp16 planes [2, B, H, W, C_stored] (hi, lo*2048) — the operand format of
    sassd_conv2d_f16x3; ``channels`` of the C_stored are meaningful, the rest are zero."""

    def __init__(self, planes, channels, tile_dist=None, reach=0, const=None):
        # Maps that descend from a scattered sparse tensor are constant over large regions.  tile_dist: int32
        # [B * tiles_y * tiles_x], pixel distance of every conv tile to the nearest active cell of the scattered map;
        # reach: number of 3x3 convs applied since; const: fp32 [channels] value of the constant region (None = 0).
        # conv2d_split uses them to skip the tiles whose output is the layer's constant (see sassd_conv2d_f16x3_occ).
        self.planes, self.channels = planes, channels
        self.tile_dist, self.reach, self.const = tile_dist, reach, const

    @property
    def shape(self):
        return (self.planes.shape[1], self.planes.shape[2], self.planes.shape[3], self.channels)

    @property
    def device(self):
        return self.planes.device

    def float(self):
        """fp32 NHWC reconstruction (hi + lo/2048), for API-compat consumers and tests."""
        return (self.planes[0].float() + self.planes[1].float() * (1.0 / 2048.0))[..., : self.channels].contiguous()

    @staticmethod
    def from_float(x):
        """Test / compat helper: split an fp32 NHWC map with the same arithmetic as the kernels."""
        B, H, W, C = x.shape
        cs = (C + 63) // 64 * 64
        hi = x.half()
        lo = ((x - hi.float()) * 2048.0).half()
        planes = torch.zeros((2, B, H, W, cs), dtype=torch.float16, device=x.device)
        planes[0, ..., :C] = hi
        planes[1, ..., :C] = lo
        return SplitMap(planes, C)


TILE_OCCUPANCY = os.environ.get("SASSD_TMA_OCC", "1") != "0"     # constant-region tile skipping in the BEV convs
CONV2D_NSPLIT_MAX_TILES_STREAM = int(os.environ.get("SASSD_NSPLIT_STREAM_TILES", "296"))   # detect_stream slots: B = 1 only (+3 % e2e)
CONV2D_NSPLIT_MAX_TILES = 1184   # B <= 4 at 200x176: measured +5 % (B=1), +15 % (B=4) on `value`, nothing beyond
CONV2D_TILE_ORDER = 0       # 1 while a latency-oriented step is captured (computed tiles first, see sassd_b200.h)
CONV2D_COUNTERS = None     # bench instrumentation: {label: int32[2] device tensor} += tiles computed, += tiles
_TILE_FAR = 1 << 20


def _tile_dist(batch, H, W, device):
    if not TILE_OCCUPANCY:
        return None
    th, tw = _lib.CONV2D_TILE_H, _lib.CONV2D_TILE_W
    return torch.full((batch * ((H + th - 1) // th) * ((W + tw - 1) // tw),), _TILE_FAR, dtype=torch.int32, device=device)


def tile_skipping_valid(H, W, reach):
    """Conditions under which the constant-region rule of sassd_conv2d_f16x3_occ is exact.  (a) tile distances are
    only recorded up to SASSD_TILE_DIST_MAX pixels, so a layer further than that from the scattered map cannot tell
    "far" from "just out of range".  (b) zero padding disturbs the constant up to reach-1 pixels from the image edge
    and the kernel exempts only the outermost tile row / column: those edge tiles must be at least that deep, which
    fails for a thin partial last tile (H % 8 or W % 16 small).  The C entry point checks the same and returns
    SASSD_ERR_UNSUPPORTED; here the layer simply falls back to computing every tile."""
    th, tw = _lib.CONV2D_TILE_H, _lib.CONV2D_TILE_W
    if reach > _lib.TILE_DIST_MAX:
        return False
    last_h = H - (H - 1) // th * th
    last_w = W - (W - 1) // tw * tw
    return min(last_h, last_w, th, tw) >= reach - 1


_CONV_CONSTS = {}


def conv_constant(x_const, cin, weight, scale, shift, relu, cout):
    """Output of a conv layer on a constant input map: fp32 [cout], obtained by running the very kernel on a
    3x3-tile map filled with the constant and reading an interior pixel, so tiles that skip the computation store
    bit-identical values.  Depends on weights only (cached; warm before CUDA-graph capture)."""
    key = (weight.data_ptr(), weight._version, None if scale is None else (scale.data_ptr(), scale._version),
           None if shift is None else (shift.data_ptr(), shift._version), bool(relu), cout, cin,
           None if x_const is None else x_const.data_ptr())
    ent = _CONV_CONSTS.get(key)
    if ent is None:
        if len(_CONV_CONSTS) >= 256:
            _CONV_CONSTS.pop(next(iter(_CONV_CONSTS)))
        h, w = 3 * _lib.CONV2D_TILE_H, 3 * _lib.CONV2D_TILE_W
        m = torch.zeros((1, h, w, cin), dtype=torch.float32, device=weight.device)
        if x_const is not None:
            m += x_const[:cin].view(1, 1, 1, cin)
        _, f = conv2d_split(SplitMap.from_float(m), weight, scale, shift, relu, cout, out_split=False, out_f32=True)
        ent = (f[0, h // 2, w // 2, :cout].clone().contiguous(), weight, scale, shift, x_const)   # keep keys alive
        _CONV_CONSTS[key] = ent
    return ent[0]


def sparse_to_bev_split(feat, coors, d_rows, C, D, H, W, batch):
    planes = torch.zeros((2, batch, H, W, D * C), dtype=torch.float16, device=feat.device)
    dist = _tile_dist(batch, H, W, feat.device)
    _call("sassd_sparse_to_bev_split", None, _ptr(feat), _ptr(coors), _ptr(d_rows), feat.shape[0], C, D, H, W, batch,
          _ptr(planes), _ptr(dist), _stream())
    return SplitMap(planes, D * C, dist)


def conv2d_split(x, weight, scale, shift, relu, cout, out_split=True, out_f32=False):
    """x: SplitMap; weight [taps, cin, cout] fp32 (packed on first use).  Returns (SplitMap | None, fp32 map | None)."""
    B, H, W, cin = x.shape
    taps = weight.shape[0]
    wp = tc_pack_cached(weight, PREC_F16X3)
    d = Conv2dDesc()
    d.batch, d.H, d.W, d.cin, d.cin_stored = B, H, W, cin, x.planes.shape[-1]
    d.cout, d.taps, d.relu = cout, taps, 1 if relu else 0
    d.tile_order = CONV2D_TILE_ORDER
    # latency-oriented steps split the 3x3 256-channel layers of small maps into half-width units (sassd_b200.h)
    tiles = B * ((H + 7) // 8) * ((W + 15) // 16)
    d.n_split = 2 if (taps == 9 and cout > 128 and tiles <= (CONV2D_NSPLIT_MAX_TILES if CONV2D_TILE_ORDER else CONV2D_NSPLIT_MAX_TILES_STREAM)) else 0
    osp = of = None
    if out_split:
        cs = (cout + 63) // 64 * 64
        # the kernel writes the BN (>= cout) columns it computes; stored channels beyond that must read as zero
        alloc = torch.empty if cs == cout else torch.zeros
        osp = alloc((2, B, H, W, cs), dtype=torch.float16, device=x.device)
        d.out_split_ch = cs
    if out_f32:
        stride = (cout + 3) // 4 * 4
        of = torch.empty((B, H, W, stride), dtype=torch.float32, device=x.device)
        d.out_f32_stride = stride
    label = "conv2d_tma[taps=%d %d->%d]" % (taps, cin, cout)
    dist = x.tile_dist if TILE_OCCUPANCY else None
    reach, cvec = 0, None
    if dist is not None:
        reach = x.reach + (1 if taps == 9 else 0)
        if not tile_skipping_valid(H, W, reach):
            dist = None        # the map is computed in full from here on (see tile_skipping_valid)
    if dist is not None:
        cvec = conv_constant(x.const, cin, weight, scale, shift, relu, cout)
    _call("sassd_conv2d_f16x3_occ", label, ctypes.byref(d), _ptr(x.planes), _ptr(wp), _ptr(scale), _ptr(shift), _ptr(of),
          _ptr(osp), _ptr(dist), reach, _ptr(cvec), _ptr(CONV2D_COUNTERS.get(label) if CONV2D_COUNTERS is not None else None),
          _stream())
    return (SplitMap(osp, cout, dist, reach, cvec) if osp is not None else None), of


# ---------------------------------------------------------------------------- sparse conv on split rows
def features_to_split(feat, d_rows=None):
    """fp32 rows [cap, C] -> split rows [2, cap, cs] fp16 (cs = C rounded up to 8)."""
    cap, C = feat.shape
    cs = (C + 7) // 8 * 8
    out = torch.empty((2, cap, cs), dtype=torch.float16, device=feat.device)
    _call("sassd_features_to_split", None, _ptr(feat), _ptr(d_rows), cap, C, cs, _ptr(out), _stream())
    return out


def split_rows_float(planes, channels):
    """fp32 reconstruction of split rows (API-compat consumers and tests)."""
    return (planes[0].float() + planes[1].float() * (1.0 / 2048.0))[:, :channels].contiguous()


SPCONV_COUNTERS = None     # bench instrumentation: int32[2] device tensor -> += executed (tile, chunk) pairs, += tiles
SPCONV_TAP_SKIP = os.environ.get("SASSD_SPS_SKIP", "1") != "0"      # use the rulebook's tile masks
SPCONV_TAP_SPLIT = os.environ.get("SASSD_SPS_SPLIT", "1") != "0"    # cluster tap split for layers with few tiles


def spconv_split(planes, weight, scale, shift, relu, cout, rows_cap, nbr=None, d_rows=None, want_f32=False,
                 tile_mask=None, ws=None):
    """planes [2, in_cap, cin_stored] fp16; weight [taps, cin, cout] fp32 (packed on first use).
    Returns (out planes [2, rows_cap, out_ch], fp32 rows or None)."""
    taps = weight.shape[0]
    wp = spconv_pack_cached(weight, planes.shape[2])
    d = SpconvDesc()
    d.cin, d.cout, d.taps = planes.shape[2], cout, taps
    d.rows_cap, d.in_rows_cap, d.relu = rows_cap, planes.shape[1], 1 if relu else 0
    d.out_ch = (cout + 7) // 8 * 8
    out = torch.empty((2, rows_cap, d.out_ch), dtype=torch.float16, device=planes.device)
    of = None
    if want_f32:
        d.out_f32_stride = (cout + 3) // 4 * 4
        of = torch.empty((rows_cap, d.out_f32_stride), dtype=torch.float32, device=planes.device)
    label = "spconv_split[taps=%d %d->%d]" % (taps, weight.shape[1], cout)
    w = None
    if SPCONV_TAP_SPLIT and taps > 1:
        w = (ws or _WS).get("spconv_split", _L().sassd_spconv_workspace_bytes(), planes.device)
    _call("sassd_spconv_f16x3", label, ctypes.byref(d), _ptr(planes), _ptr(wp), _ptr(scale), _ptr(shift), _ptr(nbr),
          _ptr(tile_mask if SPCONV_TAP_SKIP else None), _ptr(d_rows), _ptr(out), _ptr(of), _ptr(w),
          0 if w is None else w.numel(), _ptr(SPCONV_COUNTERS), _stream())
    return out, of


def split_rows_to_bev(planes, coors, d_rows, C, D, H, W, batch):
    bev = torch.zeros((2, batch, H, W, D * C), dtype=torch.float16, device=planes.device)
    dist = _tile_dist(batch, H, W, planes.device)
    _call("sassd_split_rows_to_bev", None, _ptr(planes), _ptr(coors), _ptr(d_rows), planes.shape[1], C, D, H, W, batch,
          _ptr(bev), _ptr(dist), _stream())
    return SplitMap(bev, D * C, dist)
