"""Host-side mirror of ``torchbox3d/nn/decoders/range_decoder.py``.

``RangeDecoder`` keeps the reference's five dataclass fields and ``decode`` signature
(nn/decoders/range_decoder.py:20-36), so a Hydra config swaps it in with
``_decoder._target_: rv3d.nn.decoders.range_decoder.RangeDecoder``.  Behind it the dense head
outputs go through ONE fused CUDA kernel per (stride, task) (sigmoid*mask, class max, threshold,
range-partition subsampling, fp64 box decode, compaction), then segment sort + exact greedy
rotated / weighted NMS, without the per-sweep / per-class Python loops and host syncs of the
reference (math/ops/nms.py:210-242)."""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any, Dict, Mapping, Sequence, Tuple, Union

import torch
from torch import Tensor

from ... import _native as N
from ..._pipeline import Workspace, cart_as, dtype_code, new_candidates, run_nms, threshold_as
from ..._util import ptr, require_cuda, stream_ptr

__all__ = ["RangeDecoder", "sample_by_range"]


def _mask_u8(mask: Tensor) -> Tensor:
    if mask.dtype == torch.bool:
        return mask.contiguous().view(torch.uint8)
    if mask.dtype == torch.uint8:
        return mask.contiguous()
    # the reference multiplies the scores by the mask whatever its dtype (range_decoder.py:50); the library takes the
    # validity form (the reference builds it as `range > 0`, prototype/loader.py:645-650): a 0 / 1 mask of any dtype
    return (mask != 0).contiguous().view(torch.uint8)


@dataclass
class RangeDecoder:
    enable_azimuth_invariant_targets: bool
    enable_sample_by_range: bool

    lower_bounds: Sequence[float]
    upper_bounds: Sequence[float]
    subsampling_rates: Sequence[int]

    _ws: Workspace = field(default_factory=Workspace, init=False, repr=False, compare=False)

    _cache: dict = field(default_factory=dict, init=False, repr=False, compare=False)

    # event recorded by decode_async behind the dense decode kernel (None before the first call)
    dense_done: Any = field(default=None, init=False, repr=False, compare=False)

    # The host work in front of a decode launch (partition struct, candidate counts, parameter structs) only depends on
    # the decoder's fields and the tensor shapes: it is built once per distinct key and reused, which keeps the launch
    # ahead of the rasterizer's ~77 us instead of leaving the GPU idle behind it.
    def _parts_key(self) -> Tuple:
        """The partition VALUES (cache keys must not depend on object identity: the fields are mutable)."""
        return ("parts", bool(self.enable_sample_by_range), tuple(float(v) for v in self.lower_bounds),
                tuple(float(v) for v in self.upper_bounds), tuple(int(v) for v in self.subsampling_rates))

    def _partitions(self) -> N.Partitions:
        key = self._parts_key()
        parts = self._cache.get(key)
        if parts is None:
            if not self.enable_sample_by_range:
                parts = N.make_partitions([], [], [])
            else:
                parts = N.make_partitions(list(self.lower_bounds), list(self.upper_bounds), list(self.subsampling_rates))
            self._cache[key] = parts
        return parts

    def _num_candidates(self, parts: N.Partitions, H: int, W: int) -> int:
        key = ("k", self._parts_key(), H, W)
        k = self._cache.get(key)
        if k is None:
            k = self._cache[key] = int(N.lib().rv3d_num_candidates(parts, H, W))
        return k

    def candidates(self, multiscale_outputs: Mapping[Union[int, str], Mapping[Any, Any]],
                   post_processing_config: Mapping[str, Any], task_config: Mapping[Any, Sequence[str]]):
        """Fused decode + threshold + compaction of every (stride, task).  Enqueues work only; no host sync."""
        parts = self._partitions()
        lib = N.lib()
        total_classes = sum(len(g) for g in task_config.values())
        plan, total_candidates = [], 0
        for _stride, ms in multiscale_outputs.items():                     # range_decoder.py:39
            cart, mask = ms["cart"], ms["mask"]
            B, _, H, W = cart.shape
            k = self._num_candidates(parts, H, W)
            task_offset = 0
            for task_id, group in task_config.items():                     # :45
                plan.append((ms, task_id, task_offset, total_candidates, H, W, B))
                task_offset += len(group)                                  # :78
                total_candidates += k
        if not plan:
            raise ValueError("empty multiscale_outputs / task_config")
        B = plan[0][6]
        dev = require_cuda(plan[0][0]["cart"])
        cand = new_candidates(self._ws, B, total_classes, total_candidates, dev, score_bits=N.SCORE_BITS_DECODE)
        for ms, task_id, task_offset, cand_offset, H, W, b in plan:
            if b != B:
                raise ValueError("all strides must share the batch size")
            out = ms[task_id]
            logits = out["logits"].contiguous()
            dt = logits.dtype
            reg = out["regressands"].to(dt).contiguous()
            cart = cart_as(dt, ms["cart"])
            mask = _mask_u8(ms["mask"])
            require_cuda(logits, reg, cart, mask)
            pkey = ("p", self._parts_key(), B, logits.shape[1], H, W, dt, cart.dtype, bool(self.enable_azimuth_invariant_targets),
                    task_offset, cand_offset, total_candidates, total_classes, cand.keys.numel(),
                    float(post_processing_config["min_confidence"]))
            p = self._cache.get(pkey)
            if p is None:
                p = N.DecodeParams()
                p.batch, p.n_classes, p.height, p.width = B, logits.shape[1], H, W
                p.dtype, p.cart_dtype = dtype_code(dt), dtype_code(cart.dtype)
                p.azimuth_invariant = int(bool(self.enable_azimuth_invariant_targets))
                p.category_offset, p.candidate_offset = task_offset, cand_offset
                p.total_candidates, p.total_classes = total_candidates, total_classes
                p.capacity = cand.keys.numel()
                p.min_confidence = threshold_as(dt, float(post_processing_config["min_confidence"]))
                p.parts = parts
                if len(self._cache) > 256:
                    self._cache.clear()
                self._cache[pkey] = p
            N.check(lib.rv3d_decode_compact(p, ptr(logits), ptr(reg), ptr(cart), ptr(mask), ptr(cand.keys),
                                            ptr(cand.boxes), ptr(cand.counter), stream_ptr(dev)),
                    "rv3d_decode_compact")
        return cand

    def decode_async(self, multiscale_outputs: Dict[Union[int, str], Dict[Any, Any]],
                     post_processing_config: Mapping[str, Any], task_config: Mapping[Any, Sequence[str]],
                     gather=None, stats: Tensor = None, **kwargs: Any):
        """The whole decode + NMS + pack step ENQUEUED on the current stream -> ``rv3d._pipeline.Detections`` (padded
        device buffers + the detection count on the device).  No host read, no synchronisation, fixed launch geometry:
        the call can be captured in a CUDA graph together with the rasterizer and replayed.  This is what removes the
        reference's host sync per sweep and per class (math/ops/nms.py:210-215, :22-23).

        ``gather=(PeerGather, slot, sweep_offset[, seq])``: the pack kernel also stores this rank's detections into
        every rank's gather buffer (rv3d.distributed.PeerGather)."""
        mode = str(post_processing_config["nms_mode"]).upper()
        if mode not in ("HARD", "WEIGHTED"):
            raise NotImplementedError(f"NMS Mode: {mode} is not implemented.")
        cand = self.candidates(multiscale_outputs, post_processing_config, task_config)
        # The dense (whole-GPU) part of the step is enqueued; what follows is one CTA per (sweep, class) segment.  A caller
        # with independent work for the idle SMs (the rasterizer of the next batch) makes its stream wait on this event.
        self.dense_done = torch.cuda.Event()
        self.dense_done.record()
        thr = float(post_processing_config["min_confidence"])
        score_range = (max(thr, 0.0), 1.0) if thr < 1.0 else (0.0, 0.0)     # sigmoid * mask lies in [0, 1]
        peer_kw = {}
        if gather is not None:
            peer_kw = dict(peer=gather[0], peer_slot=gather[1], sweep_offset=gather[2],
                           peer_seq=gather[3] if len(gather) > 3 else 0)
        return run_nms(self._ws, cand, post_processing_config["num_pre_nms"], post_processing_config["num_post_nms"],
                       post_processing_config["nms_threshold"], mode, N.OUT_QUAT, stats=stats, score_range=score_range,
                       **peer_kw)

    def decode(self, multiscale_outputs: Dict[Union[int, str], Dict[Any, Any]],
               post_processing_config: Mapping[str, Any], task_config: Mapping[Any, Sequence[str]],
               use_nms: bool = True, **kwargs: Any) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
        """Drop-in for RangeDecoder.decode (nn/decoders/range_decoder.py:29-124) ->
        (params (K,10) [x,y,z,l,w,h,qw,qx,qy,qz], scores (K,), categories (K,), batch_index (K,)).
        With NMS, categories / batch_index are float32 (math/ops/nms.py:51,242); without, int64.

        The returned tensors have their exact size, which only the device knows: with NMS the host waits ONCE, after
        every kernel of the step has been enqueued, on the event of this call (``decode_async`` + ``Detections.result``);
        callers that can work with padded buffers use ``decode_async`` and never wait.  The tensors are owned by the
        caller (a later call does not overwrite them).

        Multi-GPU extra (not in the reference): ``gather=(PeerGather, slot, sweep_offset)``, see ``decode_async``."""
        gather = kwargs.pop("gather", None)
        del kwargs                                                         # tools/benchmark.py passes data=
        first = next(iter(multiscale_outputs.values()))
        dt = first[next(iter(task_config.keys()))]["logits"].dtype
        if use_nms:
            return self.decode_async(multiscale_outputs, post_processing_config, task_config, gather=gather).result(dt)
        if gather is not None:
            raise ValueError("gather= needs use_nms=True (the fused gather lives in the NMS pack kernel)")
        cand = self.candidates(multiscale_outputs, post_processing_config, task_config)
        n = cand.count()
        dev = cand.keys.device
        params = torch.empty((n, 10), dtype=torch.float32, device=dev)
        scores = torch.empty((n,), dtype=torch.float32, device=dev)
        cats = torch.empty((n,), dtype=torch.int64, device=dev)
        bidx = torch.empty((n,), dtype=torch.int64, device=dev)
        lib = N.lib()
        work = self._ws.bytes("pack_scratch", lib.rv3d_pack_candidates_scratch_bytes(n), dev)
        N.check(lib.rv3d_pack_candidates(ptr(cand.keys), ptr(cand.boxes), n, cand.batch, cand.total_classes,
                                         cand.total_candidates, cand.score_bits, ptr(params), ptr(scores), ptr(cats), ptr(bidx),
                                         ptr(work), work.numel(), stream_ptr(dev)), "rv3d_pack_candidates")
        return params.to(dt), scores.to(dt), cats, bidx


def sample_by_range(scores: Tensor, categories: Tensor, cuboids: Tensor, cart: Tensor,
                    lower_bounds: Tuple[float, ...], upper_bounds: Tuple[float, ...],
                    subsampling_rates: Tuple[int, ...]) -> Tuple[Tensor, Tensor, Tensor]:
    """Drop-in for nn/decoders/range_decoder.py:127-156: scores / categories (B,1,H,W), cuboids (B,7,H,W),
    cart (B,3,H,W) -> scores (B,K), categories (B,K), cuboids (B,K,7)."""
    dev = require_cuda(scores, categories, cuboids, cart)
    B, _, H, W = cuboids.shape
    parts = N.make_partitions(list(lower_bounds), list(upper_bounds), list(subsampling_rates))
    lib = N.lib()
    K = int(lib.rv3d_num_candidates(parts, H, W))
    o_s = torch.empty((B, K), dtype=torch.float32, device=dev)
    o_c = torch.empty((B, K), dtype=torch.int64, device=dev)
    o_b = torch.empty((B, K, 7), dtype=torch.float32, device=dev)
    N.check(lib.rv3d_sample_by_range(ptr(scores.float().contiguous()), ptr(categories.to(torch.int64).contiguous()),
                                     ptr(cuboids.float().contiguous()), ptr(cart.float().contiguous()), parts, B, H, W,
                                     ptr(o_s), ptr(o_c), ptr(o_b), stream_ptr(dev)), "rv3d_sample_by_range")
    return o_s.to(scores.dtype), o_c.to(categories.dtype), o_b.to(cuboids.dtype)
