"""Shared host-side plumbing of the decode -> NMS path: compaction buffers, the NMS launch and
the two host reads of device counters.  All arithmetic happens in librv3d.so."""
from __future__ import annotations

import functools
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

from . import _native as N
from ._util import ptr, scratch, stream_ptr

_DTYPES = {torch.float32: N.F32, torch.float16: N.F16, torch.bfloat16: N.BF16}


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DTYPES[dt]
    except KeyError:
        raise TypeError(f"rv3d: unsupported dtype {dt}; use float32, float16 or bfloat16") from None


def cart_as(head_dtype: torch.dtype, cart: torch.Tensor) -> torch.Tensor:
    """cart keeps its own dtype (coding.py:128 widens it separately from the regressands; under autocast the
    heads are half precision and cart float32).  The library takes cart in the heads' dtype or in float32:
    half-precision cart next to float32 heads is widened (exact); float64 is not part of the path."""
    if cart.dtype == head_dtype or cart.dtype == torch.float32:
        return cart.contiguous()
    if cart.dtype in (torch.float16, torch.bfloat16):
        return cart.float().contiguous()
    raise TypeError(f"rv3d: unsupported cart dtype {cart.dtype}; use float32, float16 or bfloat16")


@functools.lru_cache(maxsize=64)
def threshold_as(dt: torch.dtype, value: float) -> float:
    """torch compares ``scores >= python_float`` in the tensor's dtype: round the scalar the same way.
    (Cached: building a tensor per call costs ~10 us of host time in front of every decode launch.)"""
    return float(torch.tensor(float(value), dtype=dt).float())


@dataclass
class Candidates:
    """Compaction output: rows [0, n) of keys / boxes are live."""
    keys: torch.Tensor      # (capacity,) int64 storage of the uint64 sort keys
    boxes: torch.Tensor     # (capacity, 8) f32 [x,y,z,l,w,h,yaw,score]
    counter: torch.Tensor   # (1,) i32 device counter
    batch: int
    total_classes: int
    total_candidates: int
    score_bits: int = 32    # width of the keys' score field: 31 from rv3d_decode_compact, 32 from rv3d_compact_candidates

    def count(self) -> int:
        n = int(self.counter.item())                    # host read #1 (stream sync)
        if n > self.keys.numel():
            raise N.Rv3dError(-6, "candidate compaction")
        return n


class Workspace:
    """Caches device buffers between calls (shapes repeat from batch to batch)."""

    def __init__(self) -> None:
        self._bufs: Dict[Tuple, torch.Tensor] = {}

    def get(self, tag: str, shape, dtype, device) -> torch.Tensor:
        key = (tag, tuple(shape), dtype, str(device))
        t = self._bufs.get(key)
        if t is None:
            # one live buffer per tag: drop a differently-shaped predecessor
            for k in [k for k in self._bufs if k[0] == tag and k[3] == str(device)]:
                del self._bufs[k]
            t = torch.empty(shape, dtype=dtype, device=device)
            self._bufs[key] = t
        return t

    def bytes(self, tag: str, nbytes: int, device) -> torch.Tensor:
        key = (tag, "bytes", str(device))
        t = self._bufs.get(key)
        if t is None or t.numel() < nbytes:
            t = scratch(nbytes, device)
            self._bufs[key] = t
        return t


def new_candidates(ws: Workspace, batch: int, total_classes: int, total_candidates: int, device,
                   score_bits: int = 32) -> Candidates:
    cap = batch * total_candidates
    if cap >= 2 ** 31:
        raise N.Rv3dError(N.ERR_KEYBITS, "candidate compaction (batch * candidates >= 2^31; split the batch)")
    keys = ws.get("keys", (cap,), torch.int64, device)
    boxes = ws.get("boxes", (cap, 8), torch.float32, device)
    counter = ws.get("counter", (1,), torch.int32, device)
    counter.zero_()
    return Candidates(keys, boxes, counter, batch, total_classes, total_candidates, score_bits)


_NMS_PARAMS: Dict[Tuple, "N.NmsParams"] = {}


def run_nms(ws: Workspace, cand: Candidates, n: int, num_pre_nms: int, num_post_nms: int, iou_threshold: float,
            mode: str, layout: int, merge_threshold: float = 0.5, stats: Optional[torch.Tensor] = None,
            peer=None, peer_slot: int = 0, sweep_offset: int = 0):
    """-> (params (M,10|7) f32, scores (M,) f32, categories (M,) f32, batch_index (M,) f32).
    ``peer`` (rv3d.distributed.PeerGather): also store the detections into every rank's gather buffer from inside the
    pack kernel (slot ``peer_slot``, sweep indices shifted by ``sweep_offset``)."""
    mode = mode.upper()                                                      # nms.py:207
    if mode not in ("HARD", "WEIGHTED"):
        raise NotImplementedError(f"NMS Mode: {mode} is not implemented.")   # nms.py:239-240
    dev = cand.keys.device
    S = cand.batch * cand.total_classes
    cap = min(n, S * int(num_post_nms))
    width = 10 if layout == N.OUT_QUAT else 7
    out_params = ws.get("out_params", (max(cap, 1), width), torch.float32, dev)
    out_scores = ws.get("out_scores", (max(cap, 1),), torch.float32, dev)
    out_cats = ws.get("out_cats", (max(cap, 1),), torch.float32, dev)
    out_batch = ws.get("out_batch", (max(cap, 1),), torch.float32, dev)
    out_count = ws.get("out_count", (1,), torch.int32, dev)
    # This runs right after the host has read the candidate count, i.e. with the GPU idle: the parameter block is
    # cached per configuration and only its count-dependent fields are refreshed.
    pkey = (cand.batch, cand.total_classes, cand.total_candidates, int(num_pre_nms), int(num_post_nms), mode,
            float(iou_threshold), float(merge_threshold), layout, cand.score_bits)
    p = _NMS_PARAMS.get(pkey)
    if p is None:
        p = N.NmsParams()
        p.batch, p.total_classes, p.total_candidates = cand.batch, cand.total_classes, cand.total_candidates
        p.num_pre_nms, p.num_post_nms = int(min(num_pre_nms, 2 ** 31 - 1)), int(min(num_post_nms, 2 ** 31 - 1))
        p.mode = N.NMS_HARD if mode == "HARD" else N.NMS_WEIGHTED
        # nms.py:44 hands detectron2 an f32 tensor; nms.py:101-107 hands TorchEx python floats -> C float
        p.iou_threshold = threshold_as(torch.float32, float(iou_threshold))
        p.merge_threshold = float(merge_threshold)
        p.out_layout, p.score_bits = layout, cand.score_bits
        if len(_NMS_PARAMS) > 64:
            _NMS_PARAMS.clear()
        _NMS_PARAMS[pkey] = p
    p.n_candidates, p.out_capacity = n, cap
    p.peer_world = 0
    if peer is not None:
        if layout != N.OUT_QUAT:
            raise ValueError("the fused gather carries params(10) rows (RangeDecoder.decode layout)")
        p.peer_world, p.peer_rank, p.peer_capacity, p.sweep_offset = peer.world, peer.rank, peer.capacity, int(sweep_offset)
        for q, addr in enumerate(peer.slot_ptrs(peer_slot)):
            p.peer_rows[q] = addr
    lib = N.lib()
    need = lib.rv3d_nms_scratch_bytes(p)
    work = ws.bytes("nms_scratch", need, dev)
    N.check(lib.rv3d_nms(p, ptr(cand.keys), ptr(cand.boxes), ptr(out_params), ptr(out_scores), ptr(out_cats),
                         ptr(out_batch), ptr(out_count), ptr(stats), ptr(work), work.numel(), stream_ptr(dev)),
            "rv3d_nms")
    m = int(out_count.item())                                                # host read #2 (stream sync)
    return out_params[:m], out_scores[:m], out_cats[:m], out_batch[:m]
