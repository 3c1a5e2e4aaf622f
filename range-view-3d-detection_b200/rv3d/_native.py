"""ctypes binding of librv3d.so (include/rv3d.h).  No torch types cross this boundary: only raw
device pointers, sizes, POD parameter structs and the CUDA stream handle.

The library is the ONLY implementation: there is no CPU or eager-PyTorch fallback.  If it has not
been built (``python range-view-3d-detection_b200/build.py``) every operator raises."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

LIB_PATH = Path(__file__).resolve().parent / "_lib" / "librv3d.so"
MAX_PARTITIONS = 8

RV3D_OK = 0
ERR_KEYBITS = -5
COL_LIBRARY, COL_CONVERTER, COL_CONVERTER_UNIFORM = 0, 1, 2
SCORE_BITS_DECODE = 31   # include/rv3d.h RV3D_SCORE_BITS_DECODE
F32, F16, BF16 = 0, 1, 2
NMS_HARD, NMS_WEIGHTED = 0, 1
OUT_QUAT, OUT_YAW = 0, 1


class RasterParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("max_points", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("azimuth_bins", C.c_int32), ("num_lasers", C.c_int32), ("col_mode", C.c_int32),
                ("reserved", C.c_int32), ("lidar_offset", C.c_double * 3), ("min_distance", C.c_double)]


class InputsParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("height", C.c_int32), ("width", C.c_int32), ("x_stride", C.c_int32),
                ("pad", C.c_int32), ("pad_mode", C.c_int32), ("n_features", C.c_int32),
                ("feature_channel", C.c_int32 * 8), ("tanh_channel", C.c_int32)]


class Partitions(C.Structure):
    _fields_ = [("n_partitions", C.c_int32), ("lower", C.c_float * MAX_PARTITIONS),
                ("upper", C.c_float * MAX_PARTITIONS), ("rate", C.c_int32 * MAX_PARTITIONS)]


class DecodeParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_classes", C.c_int32), ("height", C.c_int32), ("width", C.c_int32),
                ("dtype", C.c_int32), ("cart_dtype", C.c_int32), ("azimuth_invariant", C.c_int32), ("category_offset", C.c_int32),
                ("candidate_offset", C.c_int32), ("total_candidates", C.c_int32), ("total_classes", C.c_int32),
                ("capacity", C.c_int32), ("min_confidence", C.c_float), ("parts", Partitions)]


class NmsParams(C.Structure):
    _fields_ = [("batch", C.c_int32), ("total_classes", C.c_int32), ("total_candidates", C.c_int32),
                ("num_pre_nms", C.c_int32), ("num_post_nms", C.c_int32), ("mode", C.c_int32),
                ("iou_threshold", C.c_float), ("merge_threshold", C.c_float), ("capacity", C.c_int32),
                ("out_capacity", C.c_int32), ("out_layout", C.c_int32), ("score_bits", C.c_int32),
                ("score_lo", C.c_float), ("score_hi", C.c_float), ("flags", C.c_int32),
                ("peer_world", C.c_int32), ("peer_rank", C.c_int32), ("peer_capacity", C.c_int32),
                ("sweep_offset", C.c_int32), ("reserved", C.c_int32),
                ("peer_rows", C.c_void_p * 8), ("peer_seq", C.c_void_p), ("peer_slot_stride", C.c_int64),
                ("host_count", C.c_void_p)]


class Rv3dError(RuntimeError):
    def __init__(self, status: int, where: str):
        self.status = status
        super().__init__(f"{where}: {lib().rv3d_strerror(status).decode()} (rv3d status {status})")


_P, _I32, _I64, _F, _D, _SZ = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_double, C.c_size_t
_SIGNATURES = {
    "rv3d_version": (C.c_int, []),
    "rv3d_strerror": (C.c_char_p, [C.c_int]),
    "rv3d_rasterize_scratch_bytes": (_SZ, [C.POINTER(RasterParams)]),
    "rv3d_rasterize": (C.c_int, [C.POINTER(RasterParams), _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_zbuffer_scratch_bytes": (_SZ, [_I32, _I32]),
    "rv3d_zbuffer": (C.c_int, [_P, _P, _P, _I32, _P, _I32, _I32, _I64, _I32, _I32, _D, _P, _P, _P, _SZ, _P]),
    "rv3d_cart_to_sph": (C.c_int, [_P, _P, _I64, _P]),
    "rv3d_range_view_coordinates": (C.c_int, [_P, _P, _P, _I32, _I64, _I32, _I32, _I32, _P, _P]),
    "rv3d_range_view_inputs": (C.c_int, [C.POINTER(InputsParams), _P, _P, _P, _P, _P]),
    "rv3d_rasterize_inputs": (C.c_int, [C.POINTER(RasterParams), C.POINTER(InputsParams), _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_subsample_range_view": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P]),
    "rv3d_decode_range_view": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _P]),
    "rv3d_num_candidates": (_I64, [C.POINTER(Partitions), _I32, _I32]),
    "rv3d_sample_by_range": (C.c_int, [_P, _P, _P, _P, C.POINTER(Partitions), _I32, _I32, _I32, _P, _P, _P, _P]),
    "rv3d_decode_compact": (C.c_int, [C.POINTER(DecodeParams), _P, _P, _P, _P, _P, _P, _P, _P]),
    "rv3d_compact_candidates": (C.c_int, [_P, _P, _P, _I32, _I32, _I32, _F, _I32, _I32, _P, _P, _P, _P]),
    "rv3d_nms_scratch_bytes": (_SZ, [C.POINTER(NmsParams)]),
    "rv3d_nms": (C.c_int, [C.POINTER(NmsParams), _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_peer_wait": (C.c_int, [_P, _I64, _I32, _I32, _P, _P]),
    "rv3d_pair_decisions": (C.c_int, [_P, _P, _I64, _F, _I32, _P, _P, _P, _P]),
    "rv3d_debug_fastmath": (C.c_int, [_I32, _P, _P, _P, _I64, _P]),
    "rv3d_debug_column": (C.c_int, [C.POINTER(RasterParams), _P, _I64, _P, _P, _P]),
    "rv3d_nms_rotated_scratch_bytes": (_SZ, [_I32]),
    "rv3d_nms_rotated": (C.c_int, [_P, _P, _I32, _F, _P, _P, _P, _SZ, _P]),
    "rv3d_wnms_scratch_bytes": (_SZ, [_I32, _I32]),
    "rv3d_wnms": (C.c_int, [_P, _P, _I32, _I32, _F, _F, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_iou3d_aligned": (C.c_int, [_P, _P, _I64, _P, _P, _P, _P]),
    "rv3d_yaw_to_quat": (C.c_int, [_P, _P, _I64, _P]),
    "rv3d_pack_candidates_scratch_bytes": (_SZ, [_I32]),
    "rv3d_pack_candidates": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I32, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_box_iou_rotated": (C.c_int, [_P, _I64, _P, _I64, _I32, _P, _P]),
    "rv3d_classification_targets_scratch_bytes": (_SZ, [_I32, _I32, _I32, _I32, _I32]),
    "rv3d_classification_targets": (C.c_int, [_P, _P, _P, _P, _P, _P, _I32, _I32, _I32, _I32, _I32, _I32, _I32, _F, _I32,
                                              _P, _P, _P, _P, _P, _P, _SZ, _P]),
    "rv3d_instance_topk_scratch_bytes": (_SZ, [_I64]),
    "rv3d_instance_topk": (C.c_int, [_P, _P, _I64, _I32, _I32, _P, _P, _SZ, _P]),
    "rv3d_detection_records_scratch_bytes": (_SZ, [_I64]),
    "rv3d_detection_records": (C.c_int, [_P, _P, _P, _P, _I64, _P, _I32, _F, _I32, _P, _P, _P, _SZ, _P]),
    "rv3d_records_sort_unique_scratch_bytes": (_SZ, [_I64]),
    "rv3d_records_sort_unique": (C.c_int, [_P, _P, _I64, _P, _P, _P, _SZ, _P]),
    "rv3d_records_group_offsets": (C.c_int, [_P, _P, _I32, _P, _P]),
    "rv3d_unmotion_compensate": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _I32, _P, _P, _P, _P, _P]),
    "rv3d_pose_intervals": (C.c_int, [_P, _I32, _P, _P]),
    "rv3d_unmotion_compensate_table": (C.c_int, [_P, _P, _I64, _I64, _P, _P, _P, _I32, _I64, _I64, _P, _P, _P, _P, _P, _P]),
    "rv3d_transform_points": (C.c_int, [_P, _I64, _P, _P, _I32, _P, _P]),
    "rv3d_correct_laser_numbers": (C.c_int, [_P, _I64, _P, _P, _I32, _P, _P, _P]),
}

_NOT_YET_BUILT: set = set()
_lib = None


def lib() -> C.CDLL:
    """Load librv3d.so; raises (never falls back) when the CUDA library is missing."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: the rv3d operators have no CPU/PyTorch fallback. "
                "Build it with `python range-view-3d-detection_b200/build.py`.")
        handle = C.CDLL(str(LIB_PATH))
        missing = [name for name in _SIGNATURES if name not in _NOT_YET_BUILT and not hasattr(handle, name)]
        if missing:
            raise RuntimeError(f"{LIB_PATH} is stale: it does not export {missing}; rebuild it")
        for name, (res, args) in _SIGNATURES.items():
            if name in _NOT_YET_BUILT:
                continue
            fn = getattr(handle, name)
            fn.restype, fn.argtypes = res, args
        _lib = handle
    return _lib


def check(status: int, where: str) -> None:
    if status != RV3D_OK:
        raise Rv3dError(status, where)


def make_partitions(lower, upper, rates) -> Partitions:
    p = Partitions()
    n = len(rates)
    if not (len(lower) == len(upper) == n) or n > MAX_PARTITIONS:
        raise ValueError("lower_bounds / upper_bounds / subsampling_rates must have equal length <= 8")
    p.n_partitions = n
    for i in range(n):
        p.lower[i], p.upper[i], p.rate[i] = float(lower[i]), float(upper[i]), int(rates[i])
    return p
