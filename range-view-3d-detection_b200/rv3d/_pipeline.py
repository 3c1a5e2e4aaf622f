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


class Detections:
    """Result of a suppression call that has only been ENQUEUED: padded device buffers plus the detection count, which
    stays on the device (and in one int of mapped pinned host memory the pack kernel also writes) until somebody asks.
    Nothing here synchronises; ``result()`` is the one place the host waits (on an event, for this call only).

    The buffers are allocated per call, so results of earlier calls are never overwritten by later ones."""

    def __init__(self, params: torch.Tensor, scores: torch.Tensor, categories: torch.Tensor, batch_index: torch.Tensor,
                 count: torch.Tensor, host_count: torch.Tensor, stream: torch.cuda.Stream, buffer: Optional[torch.Tensor] = None):
        self.params, self.scores, self.categories, self.batch_index = params, scores, categories, batch_index
        self.count, self.host_count = count, host_count      # device (1,) i32; pinned host (1,) i32
        self.buffer = buffer                                 # the one f32 block all of the above are views of (count last)
        self.event = None
        if not torch.cuda.is_current_stream_capturing():
            self.event = torch.cuda.Event()
            self.event.record(stream)
        self._stream = stream

    def wait(self) -> int:
        """Host waits for this call (event, not a device-wide sync) and returns the number of detections."""
        if self.event is not None:
            self.event.synchronize()
        else:                          # enqueued under graph capture: the caller replays the graph, then asks
            torch.cuda.synchronize(self.params.device)
        return int(self.host_count[0])

    def result(self, dtype: Optional[torch.dtype] = None):
        """-> (params (M,10|7), scores (M,), categories (M,), batch_index (M,)), exact size; empty: the reference's
        (0, W), (0,1), (0,1), (0,1) shapes (math/ops/nms.py:250-253)."""
        m = self.wait()
        dt = dtype or self.params.dtype
        if m == 0:
            e = torch.empty((0,), dtype=dt, device=self.params.device)
            return torch.empty((0, self.params.shape[1]), dtype=dt, device=self.params.device), e.view(0, 1), e.view(0, 1), e.view(0, 1)
        return self.params[:m].to(dt), self.scores[:m].to(dt), self.categories[:m].to(dt), self.batch_index[:m].to(dt)


_NMS_PARAMS: Dict[Tuple, "N.NmsParams"] = {}


class _HostCounts:
    """Ring of mapped pinned host ints the pack kernel stores detection counts to (cudaHostAlloc memory is
    device-accessible under unified addressing; the pointer is the same on both sides)."""

    def __init__(self, slots: int = 64):
        self.buf = torch.zeros(slots, dtype=torch.int32).pin_memory()
        self.next = 0

    def take(self) -> torch.Tensor:
        i = self.next
        self.next = (i + 1) % self.buf.numel()
        return self.buf[i:i + 1]


_HOST_COUNTS: Optional[_HostCounts] = None


def _host_count_slot() -> torch.Tensor:
    global _HOST_COUNTS
    if _HOST_COUNTS is None:
        _HOST_COUNTS = _HostCounts()
    return _HOST_COUNTS.take()


def run_nms(ws: Workspace, cand: Candidates, num_pre_nms: int, num_post_nms: int, iou_threshold: float,
            mode: str, layout: int, merge_threshold: float = 0.5, stats: Optional[torch.Tensor] = None,
            peer=None, peer_slot: int = 0, sweep_offset: int = 0, peer_seq: int = 0,
            score_range: Tuple[float, float] = (0.0, 0.0)) -> Detections:
    """Enqueue score bucketing + NMS + pack on the current stream -> ``Detections`` (lazy; no host read, no sync: the
    candidate count is read on the device from ``cand.counter``).
    ``peer`` (rv3d.distributed.PeerGather): also store the detections into every rank's gather buffer from inside the
    pack kernel (slot ``peer_slot``, sweep indices shifted by ``sweep_offset``)."""
    mode = mode.upper()                                                      # nms.py:207
    if mode not in ("HARD", "WEIGHTED"):
        raise NotImplementedError(f"NMS Mode: {mode} is not implemented.")   # nms.py:239-240
    dev = cand.keys.device
    S = cand.batch * cand.total_classes
    capacity = cand.keys.numel()
    cap = max(min(capacity, S * int(min(num_post_nms, 2 ** 31 - 1))), 1)
    width = 10 if layout == N.OUT_QUAT else 7
    # one fresh allocation per call (the caching allocator makes it a pointer bump): results never alias a later call's
    buf = torch.empty((cap * (width + 3) + 1,), dtype=torch.float32, device=dev)
    out_params = buf[: cap * width].view(cap, width)
    out_scores = buf[cap * width: cap * (width + 1)]
    out_cats = buf[cap * (width + 1): cap * (width + 2)]
    out_batch = buf[cap * (width + 2): cap * (width + 3)]
    out_count = buf[cap * (width + 3):].view(torch.int32)
    host_count = _host_count_slot()
    pkey = (cand.batch, cand.total_classes, cand.total_candidates, capacity, int(num_pre_nms), int(num_post_nms), mode,
            float(iou_threshold), float(merge_threshold), layout, cand.score_bits, tuple(score_range))
    p = _NMS_PARAMS.get(pkey)
    if p is None:
        p = N.NmsParams()
        p.batch, p.total_classes, p.total_candidates = cand.batch, cand.total_classes, cand.total_candidates
        p.num_pre_nms, p.num_post_nms = int(min(num_pre_nms, 2 ** 31 - 1)), int(min(num_post_nms, 2 ** 31 - 1))
        p.mode = N.NMS_HARD if mode == "HARD" else N.NMS_WEIGHTED
        # nms.py:44 hands detectron2 an f32 tensor; nms.py:101-107 hands TorchEx python floats -> C float
        p.iou_threshold = threshold_as(torch.float32, float(iou_threshold))
        p.merge_threshold = float(merge_threshold)
        p.capacity, p.out_capacity = capacity, cap
        p.out_layout, p.score_bits = layout, cand.score_bits
        p.score_lo, p.score_hi = float(score_range[0]), float(score_range[1])
        p.flags = 0
        p.scratch_bytes = N.lib().rv3d_nms_scratch_bytes(p)
        if len(_NMS_PARAMS) > 64:
            _NMS_PARAMS.clear()
        _NMS_PARAMS[pkey] = p
    p.host_count = host_count.data_ptr()
    p.peer_world, p.peer_seq, p.peer_slot_stride = 0, None, 0
    if peer is not None:
        if layout != N.OUT_QUAT:
            raise ValueError("the fused gather carries params(10) rows (RangeDecoder.decode layout)")
        p.peer_world, p.peer_rank, p.peer_capacity, p.sweep_offset = peer.world, peer.rank, peer.capacity, int(sweep_offset)
        if peer_seq:     # sequence-flag protocol: slot and step number are read from device memory by the pack kernel
            p.peer_seq, p.peer_slot_stride = peer.seq.data_ptr(), peer.slot_bytes // 4
            for q, addr in enumerate(peer.slot_ptrs(0)):
                p.peer_rows[q] = addr
        else:
            for q, addr in enumerate(peer.slot_ptrs(peer_slot)):
                p.peer_rows[q] = addr
    lib = N.lib()
    work = ws.bytes("nms_scratch", p.scratch_bytes, dev)
    if peer is not None and peer_seq:
        peer.wait_published()     # step k - 1 has arrived everywhere before step k reuses the other slot (overlaps the decode)
    N.check(lib.rv3d_nms(p, ptr(cand.keys), ptr(cand.boxes), ptr(cand.counter), ptr(out_params), ptr(out_scores),
                         ptr(out_cats), ptr(out_batch), ptr(out_count), ptr(stats), ptr(work), work.numel(),
                         stream_ptr(dev)), "rv3d_nms")
    return Detections(out_params, out_scores, out_cats, out_batch, out_count, host_count, torch.cuda.current_stream(dev), buf)
