"""Compile and load oracle/csrc/oracle.c (gcc only).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
_SRC = _HERE / "csrc" / "oracle.c"
_OUT = _HERE / "_build" / "liboracle.so"
_lib = None


def build(force: bool = False) -> Path:
    _OUT.parent.mkdir(exist_ok=True)
    if force or not _OUT.exists() or _OUT.stat().st_mtime < _SRC.stat().st_mtime:
        cmd = [
            os.environ.get("CC", "gcc"), "-O2", "-ffp-contract=off", "-fno-fast-math",
            "-shared", "-fPIC", str(_SRC), "-o", str(_OUT), "-lm",
        ]
        subprocess.run(cmd, check=True)
    return _OUT


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(str(build()))
        c = ctypes
        _lib.orc_zbuffer.restype = None
        _lib.orc_zbuffer.argtypes = [c.c_void_p, c.c_void_p, c.c_void_p, c.c_int, c.c_void_p, c.c_int,
                                     c.c_int, c.c_int64, c.c_int, c.c_int, c.c_double, c.c_void_p, c.c_void_p]
        _lib.orc_rot_iou.restype = c.c_float
        _lib.orc_rot_iou.argtypes = [c.c_void_p, c.c_void_p, c.c_double]
        _lib.orc_rot_iou_aligned.restype = None
        _lib.orc_rot_iou_aligned.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_double, c.c_void_p]
        _lib.orc_rot_iou_aligned_mmcv.restype = None
        _lib.orc_rot_iou_aligned_mmcv.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_void_p]
        _lib.orc_nms_rotated.restype = c.c_int64
        _lib.orc_nms_rotated.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_double, c.c_double,
                                         c.c_void_p, c.c_void_p, c.c_int64]
        _lib.orc_iou_bev.restype = c.c_float
        _lib.orc_iou_bev.argtypes = [c.c_void_p, c.c_void_p]
        _lib.orc_wnms.restype = c.c_int64
        _lib.orc_wnms.argtypes = [c.c_void_p, c.c_void_p, c.c_int64, c.c_int, c.c_float, c.c_float,
                                  c.c_void_p, c.c_void_p, c.c_void_p, c.c_int64]
    return _lib
