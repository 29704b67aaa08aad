"""ctypes binding of the C ABI declared in include/sassd_b200.h.

The CUDA library is the product: there is no CPU or eager fallback.  If
libsassd_b200.so is missing or a symbol is absent this module raises — loudly —
instead of degrading."""
import ctypes
import os

from . import build as _build

c_int, c_float, c_size_t, c_void_p = ctypes.c_int, ctypes.c_float, ctypes.c_size_t, ctypes.c_void_p


class VoxelParams(ctypes.Structure):
    _fields_ = [("voxel_size", c_float * 3), ("range_min", c_float * 3), ("grid", ctypes.c_int32 * 3),
                ("max_points", ctypes.c_int32), ("max_voxels", ctypes.c_int32)]


class Conv2dDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("batch", "H", "W", "cin", "cin_stored", "cout", "taps", "relu",
                                              "out_f32_stride", "out_split_ch", "tile_order", "n_split")]


class SpconvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("cin", "cout", "taps", "rows_cap", "in_rows_cap", "relu", "out_ch",
                                              "out_f32_stride")]


class GConvDesc(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("mode", "precision", "cin", "cout", "taps", "in_stride", "out_stride",
                                              "rows_cap", "batch", "H", "W", "relu")]


GCONV_TABLE, GCONV_CONV2D, GCONV_ROWS = 0, 1, 2
PREC_FP32, PREC_TF32X3, PREC_F16X3 = 0, 1, 2
CONV2D_TILE_H, CONV2D_TILE_W = 8, 16      # SASSD_CONV2D_TILE_H / _W of the header
TILE_DIST_MAX = 9                         # SASSD_TILE_DIST_MAX
SPCONV_TILE_ROWS = 128                    # SASSD_SPCONV_TILE_ROWS

OK = 0
ERRORS = {-1: "SASSD_ERR_ARG", -2: "SASSD_ERR_LAUNCH", -3: "SASSD_ERR_WORKSPACE", -4: "SASSD_ERR_UNSUPPORTED"}
FLAGS = {1: "VOXEL_CAP", 2: "ROWS_CAP", 4: "GUIDED_CAP", 8: "NMS_CAP", 16: "HASH_FULL", 32: "DET_CAP"}

P = c_void_p
_SIGNATURES = {
    "sassd_version": (c_int, []),
    "sassd_voxelize_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sassd_voxelize": (c_int, [P, P, c_int, c_int, ctypes.POINTER(VoxelParams), c_int, P, P, P, P, c_int, P, P, P,
                               c_size_t, P]),
    "sassd_voxel_mean": (c_int, [P, P, P, c_int, c_int, P, P]),
    "sassd_anchor_mask_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sassd_anchor_mask": (c_int, [P, P, c_int, c_int, c_int, c_int, P, c_int, c_int, P, P, c_size_t, P]),
    "sassd_hash_build": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P]),
    "sassd_rulebook_subm": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P]),
    "sassd_rulebook_conv_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sassd_rulebook_conv_outputs": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, c_size_t, P]),
    "sassd_set_pdl": (c_int, [c_int]),
    "sassd_rulebook_conv_outputs_hash": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_int, P, P, c_int, P, P,
                                                 c_size_t, P]),
    "sassd_rulebook_conv_nbr": (c_int, [P, P, c_int, c_int, c_int, c_int, P, P, c_int, P, P, P]),
    "sassd_rulebook_pairs": (c_int, [P, P, c_int, P, P, P]),
    "sassd_gconv": (c_int, [ctypes.POINTER(GConvDesc), P, P, P, P, P, P, P, P]),
    "sassd_gconv_pack_bytes": (c_size_t, [c_int, c_int, c_int, c_int]),
    "sassd_gconv_pack": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "sassd_conv2d_f16x3": (c_int, [ctypes.POINTER(Conv2dDesc), P, P, P, P, P, P, P]),
    "sassd_conv2d_f16x3_occ": (c_int, [ctypes.POINTER(Conv2dDesc), P, P, P, P, P, P, P, c_int, P, P, P]),
    "sassd_rotate_overlap_eval": (c_int, [P, P, P, P, P, c_int, c_int, c_int, P, P]),
    "sassd_kitti_match": (c_int, [c_int, P, P, P, P, P, P, P, P, P, P, P, P, c_int, ctypes.c_double, c_int, c_int, P, P,
                                  P, P]),
    "sassd_spconv_pack_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sassd_spconv_pack": (c_int, [P, c_int, c_int, c_int, c_int, P, P]),
    "sassd_spconv_workspace_bytes": (c_size_t, []),
    "sassd_spconv_f16x3": (c_int, [ctypes.POINTER(SpconvDesc), P, P, P, P, P, P, P, P, P, P, c_size_t, P, P]),
    "sassd_features_to_split": (c_int, [P, P, c_int, c_int, c_int, P, P]),
    "sassd_split_rows_to_bev": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "sassd_sparse_to_bev_split": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "sassd_sparse_to_bev": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P]),
    "sassd_decode_select_workspace_bytes": (c_size_t, [c_int, c_int]),
    "sassd_decode_select": (c_int, [P, c_int, c_int, c_int, c_int, c_int, P, c_int, P, c_int, c_float, P, P, P, P, c_int, P,
                                    P, c_size_t, P]),
    "sassd_pswarp": (c_int, [P, c_int, c_int, c_int, c_int, P, P, c_int, c_float, c_float, c_float, P, P]),
    "sassd_rescore_nms_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "sassd_rescore_nms": (c_int, [P, P, P, P, c_int, c_int, c_float, c_float, c_int, P, P, c_int, P, P, c_size_t, P]),
    "sassd_nms_workspace_bytes": (c_size_t, [c_int]),
    "sassd_nms_mask": (c_int, [P, c_int, c_float, P, P]),
    "sassd_nms_sorted": (c_int, [P, c_int, c_float, P, P, P, c_size_t, P]),
    "sassd_boxes_iou_bev": (c_int, [P, c_int, P, c_int, P, P]),
}

_LIB = None


class SassdError(RuntimeError):
    pass


def exported_symbols():
    return sorted(_SIGNATURES)


def load(build_if_missing=True):
    """Load libsassd_b200.so and bind every symbol of include/sassd_b200.h."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = _build.LIB
    if not os.path.exists(path):
        if not build_if_missing:
            raise SassdError("libsassd_b200.so not built: run `python -m sassd_b200.build`")
        path = _build.build()
    lib = ctypes.CDLL(path)
    for name, (res, args) in _SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise SassdError("libsassd_b200.so lacks symbol %s (stale build?)" % name) from e
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def check(rc, what):
    if rc != OK:
        raise SassdError("%s failed: %s (%d)" % (what, ERRORS.get(rc, "unknown"), rc))


def decode_flags(word):
    return [name for bit, name in FLAGS.items() if word & bit]
