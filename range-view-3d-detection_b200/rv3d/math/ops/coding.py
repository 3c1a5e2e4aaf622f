"""Host-side mirror of ``torchbox3d/math/ops/coding.py``: ``decode_range_view`` and the detection wire format
(``build_dataframe`` over plain columns instead of a polars frame)."""
from __future__ import annotations

from typing import Any, Dict, Mapping, Optional, Sequence

import numpy as np
import torch
from torch import Tensor

from ... import _native as N
from ..._pipeline import cart_as, dtype_code
from ..._util import ptr, require_cuda, scratch, stream_ptr

__all__ = ["decode_range_view", "build_records", "build_records_device", "prepare_for_evaluation", "group_by_sweep",
           "build_dataframe", "RECORD_DTYPE", "SCHEMA"]

# one detection on the wire (include/rv3d.h rv3d_detection_record, 64 bytes); the fields are the numeric columns of
# SERIALIZED_SCHEMA (nn/arch/detector.py:45-60) plus batch_index and the range the evaluation filters on
RECORD_DTYPE = np.dtype([("tx_m", "<f4"), ("ty_m", "<f4"), ("tz_m", "<f4"), ("length_m", "<f4"), ("width_m", "<f4"),
                         ("height_m", "<f4"), ("qw", "<f4"), ("qx", "<f4"), ("qy", "<f4"), ("qz", "<f4"), ("score", "<f4"),
                         ("category_index", "<i4"), ("timestamp_ns", "<i8"), ("batch_index", "<i4"), ("range_m", "<f4")])
assert RECORD_DTYPE.itemsize == 64

# column -> numpy dtype of build_dataframe's result (coding.py:11-28 with the joins of :60-76 applied)
SCHEMA = {"tx_m": np.float32, "ty_m": np.float32, "tz_m": np.float32, "length_m": np.float32, "width_m": np.float32,
          "height_m": np.float32, "qw": np.float32, "qx": np.float32, "qy": np.float32, "qz": np.float32,
          "score": np.float32, "batch_index": np.int32, "log_id": object, "timestamp_ns": np.int64, "category": object}


def decode_range_view(regressands: Tensor, cart: Tensor, enable_azimuth_invariant_targets: bool) -> Tensor:
    """Drop-in for math/ops/coding.py:110-144: (B,8,H,W), (B,3,H,W) -> (B,7,H,W) in the regressands'
    dtype; cart is read in its own dtype; fp64 arithmetic inside, one cast at the end."""
    dev = require_cuda(regressands, cart)
    if regressands.dim() != 4 or regressands.shape[1] != 8 or cart.dim() != 4 or cart.shape[1] != 3:
        raise ValueError("decode_range_view expects regressands (B,8,H,W) and cart (B,3,H,W)")
    B, _, H, W = regressands.shape
    reg = regressands.contiguous()
    crt = cart_as(reg.dtype, cart)
    out = torch.empty((B, 7, H, W), dtype=reg.dtype, device=dev)
    N.check(N.lib().rv3d_decode_range_view(ptr(reg), ptr(crt), ptr(out), dtype_code(reg.dtype), dtype_code(crt.dtype), B, H, W,
                                           int(bool(enable_azimuth_invariant_targets)), stream_ptr(dev)),
            "rv3d_decode_range_view")
    return out


def build_records_device(params: Tensor, scores: Tensor, categories: Tensor, batch_index: Tensor,
                         timestamps_ns: Optional[Sequence[int]] = None, max_range_m: Optional[float] = None):
    """The decoder's four outputs -> (records (capacity, 64) uint8 CUDA tensor of ``RECORD_DTYPE`` rows in the decoder's
    order, count (1,) int32 CUDA tensor): packed and, with ``max_range_m``, range-filtered like detector.py:573-581 on the
    device.  Enqueues work only."""
    dev = require_cuda(params)
    n = params.shape[0]
    if params.dim() != 2 or params.shape[1] != 10:
        raise ValueError("params must be (N,10) [tx,ty,tz,l,w,h,qw,qx,qy,qz]")
    p = params.float().contiguous()
    sc = scores.float().reshape(-1).contiguous()
    ca = categories.float().reshape(-1).contiguous()
    bi = batch_index.float().reshape(-1).contiguous()
    if not (sc.numel() == ca.numel() == bi.numel() == n):
        raise ValueError("params, scores, categories and batch_index disagree in length")
    stamps = None
    if timestamps_ns is not None:
        stamps = torch.as_tensor(np.asarray(timestamps_ns, dtype=np.int64), device=dev)
    out = torch.empty((max(n, 1), 64), dtype=torch.uint8, device=dev)
    count = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = N.lib()
    work = scratch(lib.rv3d_detection_records_scratch_bytes(n), dev)
    N.check(lib.rv3d_detection_records(ptr(p), ptr(sc), ptr(ca), ptr(bi), n, ptr(stamps), 0 if stamps is None else stamps.numel(),
                                       float(max_range_m or 0.0), int(max_range_m is not None), ptr(out), ptr(count), ptr(work),
                                       work.numel(), stream_ptr(dev)), "rv3d_detection_records")
    return out[:max(n, 0)] if n else out[:0], count


def build_records(params: Tensor, scores: Tensor, categories: Tensor, batch_index: Tensor,
                  timestamps_ns: Optional[Sequence[int]] = None, max_range_m: Optional[float] = None) -> np.ndarray:
    """The decoder's four outputs -> a numpy structured array of ``RECORD_DTYPE`` rows in the decoder's order, packed
    (and, with ``max_range_m``, range-filtered like detector.py:573-581) on the device and brought over in ONE copy;
    the reference reads thirteen columns back one ``.tolist()`` at a time (coding.py:42-57)."""
    rec, count = build_records_device(params, scores, categories, batch_index, timestamps_ns, max_range_m)
    m = int(count.item())
    return rec[:m].cpu().numpy().view(RECORD_DTYPE).reshape(-1)


def prepare_for_evaluation(params: Tensor, scores: Tensor, categories: Tensor, batch_index: Tensor,
                           timestamps_ns: Optional[Sequence[int]], max_range_m: float) -> np.ndarray:
    """The detection half of ``prepare_for_evaluation`` (nn/arch/detector.py:573-584) without leaving the device: range
    filter ``||(tx,ty,tz)|| <= max_range_m``, ``.sort(col("score"), descending=True)``, ``.unique()`` -> the distinct rows
    (all fields compared bit for bit) as ``RECORD_DTYPE``, score descending; ONE copy to the host.  (polars leaves the
    order of rows with equal scores, and the order after ``unique()``, unspecified; the evaluation consumes the set.)"""
    rec, count = build_records_device(params, scores, categories, batch_index, timestamps_ns, max_range_m)
    dev = rec.device
    cap = rec.shape[0]
    out = torch.empty((max(cap, 1), 64), dtype=torch.uint8, device=dev)
    out_count = torch.zeros(1, dtype=torch.int32, device=dev)
    lib = N.lib()
    work = scratch(lib.rv3d_records_sort_unique_scratch_bytes(cap), dev)
    N.check(lib.rv3d_records_sort_unique(ptr(rec) if cap else None, ptr(count), cap, ptr(out), ptr(out_count), ptr(work), work.numel(),
                                         stream_ptr(dev)), "rv3d_records_sort_unique")
    m = int(out_count.item())
    return out[:m].cpu().numpy().view(RECORD_DTYPE).reshape(-1)


def group_by_sweep(records: Tensor, count: Tensor, batch: int) -> Tensor:
    """``dts.group_by(["log_id", "timestamp_ns"], maintain_order=True)`` (nn/arch/detector.py:366-380) on a device record
    stream in the decoder's order: -> offsets (batch + 1,) int32 CUDA tensor, sweep b's records are
    ``records[offsets[b]:offsets[b + 1]]`` (each slice is what the reference writes to ``<log_id>/<timestamp_ns>.feather``)."""
    dev = require_cuda(records, count)
    offsets = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    if records.shape[0] == 0:
        return offsets.zero_()
    N.check(N.lib().rv3d_records_group_offsets(ptr(records), ptr(count), int(batch), ptr(offsets), stream_ptr(dev)),
            "rv3d_records_group_offsets")
    return offsets


def build_dataframe(params: Tensor, scores: Tensor, categories: Tensor, batch_index: Tensor, uuids: Mapping[str, Sequence[Any]],
                    idx_to_category: Sequence[str]) -> Dict[str, np.ndarray]:
    """coding.py:31-76 over plain columns: ``uuids`` = {"batch_index", "log_id", "timestamp_ns"} (one entry per sweep),
    ``idx_to_category`` = class names in index order.  -> the joined frame as a dict of numpy columns in ``SCHEMA`` order;
    like the reference's inner joins, detections whose sweep or class has no entry are dropped and the order is kept."""
    sweeps = {int(b): (str(l), int(t)) for b, l, t in zip(uuids["batch_index"], uuids["log_id"], uuids["timestamp_ns"])}
    size = (max(sweeps) + 1) if sweeps else 0
    stamps = [sweeps.get(b, ("", 0))[1] for b in range(size)]
    rec = build_records(params, scores, categories, batch_index, stamps if size else None)
    keep = np.array([int(b) in sweeps and 0 <= int(c) < len(idx_to_category)
                     for b, c in zip(rec["batch_index"], rec["category_index"])], dtype=bool)
    rec = rec[keep]
    out: Dict[str, np.ndarray] = {k: np.ascontiguousarray(rec[k]) for k in
                                  ("tx_m", "ty_m", "tz_m", "length_m", "width_m", "height_m", "qw", "qx", "qy", "qz", "score",
                                   "batch_index")}
    out["log_id"] = np.array([sweeps[int(b)][0] for b in rec["batch_index"]], dtype=object)
    out["timestamp_ns"] = np.ascontiguousarray(rec["timestamp_ns"])
    out["category"] = np.array([idx_to_category[int(c)] for c in rec["category_index"]], dtype=object)
    return out
