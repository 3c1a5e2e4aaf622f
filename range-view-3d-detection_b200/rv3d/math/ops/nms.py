"""Host-side mirror of ``torchbox3d/math/ops/nms.py``: same names, arguments and error behaviour;
the arithmetic runs in librv3d.so (segment sort + exact greedy rotated NMS / weighted NMS)."""
from __future__ import annotations

from typing import Optional, Tuple

import torch
from torch import Tensor

from ... import _native as N
from ..._pipeline import Candidates, Workspace, new_candidates, run_nms, threshold_as
from ..._util import ptr, require_cuda, scratch, stream_ptr

__all__ = ["batched_multiclass_nms", "hard_multiclass_nms", "weighted_multiclass_nms", "weighted_nms",
           "nms_rotated"]

_WS = Workspace()


def _compact(cuboids: Tensor, scores: Tensor, categories: Tensor, min_confidence: Optional[float],
             total_classes: Optional[int] = None) -> Candidates:
    dev = require_cuda(cuboids, scores, categories)
    B, K, P = cuboids.shape
    if P != 7:
        raise ValueError("cuboids must be (B,K,7) [x,y,z,l,w,h,yaw]")
    cats = categories.reshape(B, K).to(torch.int64).contiguous()
    if total_classes is None:
        # the reference discovers the classes with torch.unique (nms.py:22), i.e. a host read; callers that know the
        # class count pass `total_classes` and stay sync-free
        lo, hi = (int(v) for v in torch.stack([cats.min(), cats.max()]).tolist()) if cats.numel() else (0, 0)
        if lo < 0:
            raise ValueError("negative category index")
        total_classes = hi + 1
    cand = new_candidates(_WS, B, total_classes, max(K, 1), dev)
    sc = scores.reshape(B, K)
    thr = 0.0 if min_confidence is None else threshold_as(sc.dtype, min_confidence)
    N.check(N.lib().rv3d_compact_candidates(ptr(cuboids.float().contiguous()), ptr(sc.float().contiguous()), ptr(cats),
                                            B, K, total_classes, thr, int(min_confidence is not None),
                                            cand.keys.numel(), ptr(cand.keys), ptr(cand.boxes), ptr(cand.counter),
                                            stream_ptr(dev)), "rv3d_compact_candidates")
    return cand


def batched_multiclass_nms(cuboids: Tensor, scores: Tensor, categories: Tensor, num_pre_nms: int, num_post_nms: int,
                           iou_threshold: float, min_confidence: float, nms_mode: str
                           ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Drop-in for math/ops/nms.py:181-266.  (B,K,7),(B,K),(B,K) -> flat (M,7),(M,),(M,),(M,) ordered by
    sweep asc, class asc, score desc; categories / batch index come back float32 like upstream."""
    mode = nms_mode.upper()
    if mode not in ("HARD", "WEIGHTED"):
        raise NotImplementedError(f"NMS Mode: {mode} is not implemented.")
    cand = _compact(cuboids, scores, categories, min_confidence)
    det = run_nms(_WS, cand, num_pre_nms, num_post_nms, iou_threshold, mode, N.OUT_YAW)
    if det.wait() == 0:                                                       # nms.py:250-253
        return (cuboids.new_empty((0, cuboids.shape[-1])), scores.new_empty((0, 1)),
                categories.new_empty((0, 1)), categories.new_empty((0, 1)))
    p, s, c, b = det.result()
    return p.to(cuboids.dtype), s.to(scores.dtype), c.to(scores.dtype), b.to(scores.dtype)


def _per_sweep(cuboids_i, scores_i, categories_i, iou_threshold, num_pre_nms, num_post_nms, mode):
    if cuboids_i.shape[0] == 0:
        return cuboids_i.new_empty((0, 7)), scores_i.new_empty((0,)), scores_i.new_empty((0,))
    cand = _compact(cuboids_i[None], scores_i[None], categories_i[None], None)
    det = run_nms(_WS, cand, num_pre_nms, num_post_nms, iou_threshold, mode, N.OUT_YAW)
    m = det.wait()
    return det.params[:m].to(cuboids_i.dtype), det.scores[:m].to(scores_i.dtype), det.categories[:m].to(scores_i.dtype)


def hard_multiclass_nms(cuboids_i: Tensor, scores_i: Tensor, categories_i: Tensor, iou_threshold: float,
                        num_pre_nms: int, num_post_nms: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Drop-in for math/ops/nms.py:11-61 (one sweep, already confidence-filtered)."""
    return _per_sweep(cuboids_i, scores_i, categories_i, iou_threshold, num_pre_nms, num_post_nms, "HARD")


def weighted_multiclass_nms(cuboids_i: Tensor, scores_i: Tensor, categories_i: Tensor, iou_threshold: float,
                            num_pre_nms: int, num_post_nms: int) -> Tuple[Tensor, Tensor, Tensor]:
    """Drop-in for math/ops/nms.py:64-123 (merge_thresh = 0.5, :106)."""
    return _per_sweep(cuboids_i, scores_i, categories_i, iou_threshold, num_pre_nms, num_post_nms, "WEIGHTED")


def weighted_nms(boxes: Tensor, data2merge: Tensor, scores: Tensor, nms_threshold: float, merge_thresh: float
                 ) -> Tuple[Tensor, Tensor, Tensor]:
    """Drop-in for math/ops/nms.py:126-177 (TorchEx wnms_gpu behind it).  boxes (N,5) [x1,y1,x2,y2,ry],
    data2merge (N,C), scores (N,) -> keep (M,) original indices, output (M,C+1) merged rows with the kept
    score last, count (M,) merge-set sizes."""
    dev = require_cuda(boxes, data2merge, scores)
    n = boxes.shape[0]
    sorted_scores, order = scores.sort(dim=0, descending=True, stable=True)
    b = boxes[order].contiguous().float()
    ds = torch.cat([data2merge[order].float(), sorted_scores[:, None].float()], 1).contiguous()
    D = ds.shape[1]
    output = torch.empty_like(ds)
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    count = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    n_out = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = N.lib()
    work = scratch(lib.rv3d_wnms_scratch_bytes(n, D), dev)
    N.check(lib.rv3d_wnms(ptr(b), ptr(ds), n, D, float(nms_threshold), float(merge_thresh), ptr(output), ptr(keep),
                          ptr(count), ptr(n_out), ptr(work), work.numel(), stream_ptr(dev)), "rv3d_wnms")
    m = int(n_out.item())
    assert output[m:, :].sum() == 0 and (count[:m] > 0).all()                 # nms.py:173-174
    return order[keep[:m]].contiguous(), output[:m], count[:m]


def nms_rotated(boxes: Tensor, scores: Tensor, iou_threshold) -> Tensor:
    """Stand-in for ``detectron2.layers.nms.nms_rotated`` (call site math/ops/nms.py:41-45):
    boxes (N,5) f32 (xc,yc,w,h,angle_degrees) -> kept original indices in score order (int64)."""
    dev = require_cuda(boxes, scores)
    n = boxes.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=dev)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=dev)
    lib = N.lib()
    work = scratch(lib.rv3d_nms_rotated_scratch_bytes(n), dev)
    thr = float(torch.as_tensor(iou_threshold, dtype=torch.float32))
    N.check(lib.rv3d_nms_rotated(ptr(boxes.float().contiguous()), ptr(scores.float().contiguous()), n, thr, ptr(keep),
                                 ptr(n_keep), ptr(work), work.numel(), stream_ptr(dev)), "rv3d_nms_rotated")
    return keep[: int(n_keep.item())]
