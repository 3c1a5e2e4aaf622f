"""Host-side mirror of ``torchbox3d/math/ops/assignment.py`` (training-time callers of the path's operators,
SURVEY 8f row 4).  Same names, arguments and results; the per-instance Python loop of
``compute_classification_targets`` (:121-139) becomes one segmented top-k on the device."""
from __future__ import annotations

import math
from typing import Any, Mapping, Tuple

import torch
import torch.nn.functional as F
from torch import Tensor

from ... import _native as N
from ..._util import ptr, require_cuda, scratch, stream_ptr
from .coding import decode_range_view
from .iou import iou_3d_axis_aligned as _iou_3d_pair

__all__ = ["box_iou_rotated", "iou_2d_axis_aligned", "iou_3d_axis_aligned", "compute_classification_targets"]

XYLWA_INDICES = (0, 1, 3, 4, 6)


def box_iou_rotated(bboxes1: Tensor, bboxes2: Tensor, aligned: bool = False) -> Tensor:
    """Stand-in for ``mmcv.ops.box_iou_rotated`` as this code base calls it (mode 'iou', clockwise angles in
    radians): (N,5), (M,5) float32 -> (N,M), or (N,) when ``aligned``."""
    dev = require_cuda(bboxes1, bboxes2)
    a, b = bboxes1.float().contiguous(), bboxes2.float().contiguous()
    n, m = a.shape[0], b.shape[0]
    if a.dim() != 2 or b.dim() != 2 or a.shape[1] != 5 or b.shape[1] != 5:
        raise ValueError("box_iou_rotated expects (N,5) and (M,5) boxes")
    if aligned and n != m:
        raise ValueError("aligned IoU needs the same number of boxes on both sides")
    out = torch.empty((n,) if aligned else (n, m), dtype=torch.float32, device=dev)
    N.check(N.lib().rv3d_box_iou_rotated(ptr(a), n, ptr(b), m, int(bool(aligned)), ptr(out), stream_ptr(dev)),
            "rv3d_box_iou_rotated")
    return out


def iou_2d_axis_aligned(cuboids_a: Tensor, cuboids_b: Tensor, **kwargs: Any) -> Tensor:
    """assignment.py:64-73: clamped aligned BEV IoU.  With ``normalize_affinities`` the reference reads a name it
    never assigned (:71) -- the same error is raised here."""
    iou_bev = box_iou_rotated(cuboids_a[:, XYLWA_INDICES], cuboids_b[:, XYLWA_INDICES], aligned=True).clamp(0.0, 1.0)
    if kwargs["normalize_affinities"]:
        raise UnboundLocalError("cannot access local variable 'object_ious' where it is not associated with a value")
    return iou_bev


def iou_3d_axis_aligned(cuboids_a: Tensor, cuboids_b: Tensor, **kwargs: Any) -> Tensor:
    """assignment.py:20-61: aligned 3D IoU, optionally divided by (its maximum + 1e-8)."""
    object_ious, _ = _iou_3d_pair(cuboids_a, cuboids_b)
    if kwargs["normalize_affinities"]:
        object_ious /= object_ious.max() + 1e-8
    return object_ious


def _segment_min(values: Tensor, seg: Tensor, n_seg: int) -> Tensor:
    out = torch.full((n_seg,), float("inf"), dtype=values.dtype, device=values.device)
    return out.scatter_reduce(0, seg.long(), values, reduce="amin")


TOPK_ALL = 0x7FFFFFFF      # include/rv3d.h RV3D_TOPK_ALL


def _k_slots(k) -> int:
    """targets_config.k -> slots per instance: the production value is ``.inf`` (conf/model/range_view.yaml:126:
    ``min(k, len)`` then keeps every pixel of the instance)."""
    if isinstance(k, float) and (math.isinf(k) or k >= TOPK_ALL):
        return TOPK_ALL
    return min(int(k), TOPK_ALL)


def compute_classification_targets(input: Tensor, target: Tensor, classification_labels: Tensor, cart: Tensor,
                                   targets_config: Mapping[str, Any], mask: Tensor, panoptics: Tensor,
                                   background_index: int, max_instances: int = None) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Drop-in for assignment.py:76-148 -> (affinities (B,C,H,W), foreground_mask (B,1,H,W),
    background_mask (B,1,H,W) bool, regression_weights (B,1,H,W) bool).

    float32 tensors with k <= 64 (or k = inf, the production setting) run as ONE fused call
    (``rv3d_classification_targets``: only foreground pixels are decoded, per-instance top-k through atomic slot lists,
    no dense intermediates).  ``max_instances`` (an upper bound on the panoptic ids, exclusive) lets the caller skip the
    one host read the fused form needs to size its slot table; without it the largest id is read back.  The bound is a
    promise: an id at or above it traps the kernel (a sticky CUDA launch failure) rather than yielding partial targets.  Other dtypes,
    larger finite k and GAUSSIAN + normalize_affinities take the composed form below."""
    dev = require_cuda(input, target, cart)
    cfg = dict(targets_config)
    name = str(cfg["affinity_fn"]).upper()
    if name not in ("BEV", "GAUSSIAN"):
        raise NotImplementedError("This affinity function is not implemented.")
    k = _k_slots(cfg["k"])
    fused = (input.dtype == target.dtype == cart.dtype == torch.float32 and (k <= 64 or k == TOPK_ALL) and k >= 0
             and not (name == "GAUSSIAN" and cfg["normalize_affinities"]) and panoptics.numel() > 0)
    if not fused:
        return _compute_classification_targets_composed(input, target, classification_labels, cart, cfg, mask, panoptics,
                                                        background_index, name, k)
    if name == "BEV" and cfg["normalize_affinities"]:      # the reference's own failure mode (:71)
        if bool((panoptics > 0).any()):
            raise UnboundLocalError("cannot access local variable 'object_ious' where it is not associated with a value")
    B, _, H, W = target.shape
    C = int(background_index)
    cap = int(max_instances) if max_instances is not None else int(panoptics.max()) + 1
    cap = max(cap, 1)
    lib = N.lib()
    inp, tgt, crt = input.detach().contiguous(), target.contiguous(), cart.contiguous()
    lab = classification_labels.reshape(B, H, W).to(torch.int64).contiguous()
    pan = panoptics.reshape(B, H, W).to(torch.int64).contiguous()
    msk = mask.reshape(B, H, W).to(torch.bool).contiguous()
    affinities = torch.empty((B, C, H, W), dtype=torch.float32, device=dev)
    foreground = torch.empty((B, 1, H, W), dtype=torch.float32, device=dev)
    background = torch.empty((B, 1, H, W), dtype=torch.bool, device=dev)
    reg_w = torch.empty((B, 1, H, W), dtype=torch.bool, device=dev)
    work = scratch(lib.rv3d_classification_targets_scratch_bytes(B, H, W, k, cap), dev)
    sigma2 = float(cfg.get("sigma", 1.0)) ** 2 if name == "GAUSSIAN" else 1.0
    # status = NULL: the capacity is exact (read back above) or the caller's promise -- a broken promise traps the kernel
    N.check(lib.rv3d_classification_targets(ptr(inp), ptr(tgt), ptr(lab), ptr(crt), ptr(msk), ptr(pan), B, C, H, W,
                                            1 if name == "GAUSSIAN" else 0, int(bool(cfg["enable_azimuth_invariant_targets"])),
                                            k, sigma2, cap, ptr(affinities), ptr(foreground), ptr(background), ptr(reg_w),
                                            None, ptr(work), work.numel(), stream_ptr(dev)), "rv3d_classification_targets")
    return affinities, foreground, background, reg_w


def _compute_classification_targets_composed(input, target, classification_labels, cart, cfg, mask, panoptics,
                                             background_index, name, k):
    """The same function composed from the free-standing operators (dense decodes, gathers, segmented top-k by one
    stable sort): any floating dtype, any k."""
    dev = input.device
    all_foreground = F.one_hot(classification_labels, background_index + 1).permute(0, 3, 1, 2)[:, :-1].float()   # :91-95
    pds = decode_range_view(input.detach(), cart, True)                                                           # :105-109
    gts = decode_range_view(target, cart, bool(cfg["enable_azimuth_invariant_targets"]))                          # :110-114
    B, _, H, W = target.shape
    ids = panoptics.reshape(B, H, W)
    fg = ids > 0                                                       # one_hot(...)[:, 1:]: id 0 is background (:118)
    affinities = torch.zeros_like(target[:, 0:1])
    foreground_mask = torch.zeros_like(target[:, 0:1])
    if bool(fg.any()):
        b_idx, h_idx, w_idx = fg.nonzero(as_tuple=True)               # raster order inside a sweep == masked_select (:124-125)
        n_ids = int(ids.max()) + 1
        seg = (b_idx * n_ids + ids[fg]).to(torch.int32).contiguous()
        dts = pds.permute(0, 2, 3, 1)[fg]
        gt = gts.permute(0, 2, 3, 1)[fg]
        if name == "BEV":
            aff = iou_2d_axis_aligned(dts, gt, **cfg)                                                             # :127
        else:
            dists = torch.linalg.norm(dts[:, :3] - gt[:, :3], dim=-1)                                             # :151-159
            if cfg["normalize_affinities"]:
                dists = dists - _segment_min(dists, seg, B * n_ids)[seg.long()]
            aff = torch.exp(-dists / cfg["sigma"] ** 2)
        aff = aff.float().contiguous()
        like = torch.empty_like(aff)
        lib = N.lib()
        work = scratch(lib.rv3d_instance_topk_scratch_bytes(aff.numel()), dev)
        N.check(lib.rv3d_instance_topk(ptr(aff), ptr(seg), aff.numel(), B * n_ids, k, ptr(like), ptr(work),
                                       work.numel(), stream_ptr(dev)), "rv3d_instance_topk")
        affinities[b_idx, 0, h_idx, w_idx] = like.type_as(affinities)                                            # :134-136
        foreground_mask[b_idx, 0, h_idx, w_idx] = like.bool().type_as(affinities)                                # :137-139
    background_mask = torch.logical_and(foreground_mask.logical_not(), mask)                                     # :141
    affinities = affinities * all_foreground
    regression_weights = all_foreground.any(dim=1, keepdim=True)
    return affinities, foreground_mask, background_mask, regression_weights
