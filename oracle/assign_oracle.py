"""CPU restatement of the training-time callers.  TEST INFRASTRUCTURE ONLY.

math/ops/assignment.py (paths relative to /root/reference) with mmcv's box_iou_rotated replaced by the oracle's C
routine (orc_rot_iou, radians) -> the IoU arithmetic itself stays "parity unpinned" (un-vendored mmcv), the control
flow is pinned by tests/golden/assign.npz (verbatim reference module, same substitution)."""
from __future__ import annotations

import numpy as np
import torch

from .rv_oracle import decode_range_view, rot_iou_pairs

XYLWA = [0, 1, 3, 4, 6]


def box_iou_rotated(a: torch.Tensor, b: torch.Tensor, aligned: bool = False) -> torch.Tensor:
    a, b = a.float().numpy(), b.float().numpy()
    if aligned:
        return torch.from_numpy(rot_iou_pairs(a, b, 1.0))
    n, m = len(a), len(b)
    if n == 0 or m == 0:
        return torch.zeros((n, m))
    return torch.from_numpy(rot_iou_pairs(np.repeat(a, m, axis=0), np.tile(b, (n, 1)), 1.0).reshape(n, m))


def iou_2d_axis_aligned(a, b, **kw):                                   # :64-73
    iou = box_iou_rotated(a[:, XYLWA].contiguous(), b[:, XYLWA].contiguous(), aligned=True).clamp(0.0, 1.0)
    if kw["normalize_affinities"]:
        raise UnboundLocalError("object_ious")
    return iou


def gaussian(a, b, **kw):                                              # :151-159
    d = torch.linalg.norm(a[:, :3] - b[:, :3], dim=-1)
    if kw["normalize_affinities"]:
        d = d - d.min()
    return torch.exp(-d / kw["sigma"] ** 2)


def compute_classification_targets(inp, target, labels, cart, cfg, mask, panoptics, background_index):
    """:76-148, instance by instance."""
    onehot = torch.nn.functional.one_hot(labels, background_index + 1).permute(0, 3, 1, 2)[:, :-1].float()
    fn = {"BEV": iou_2d_axis_aligned, "GAUSSIAN": gaussian}[str(cfg["affinity_fn"]).upper()]
    pds = decode_range_view(inp.detach(), cart, True)
    gts = decode_range_view(target, cart, bool(cfg["enable_azimuth_invariant_targets"]))
    aff = torch.zeros_like(target[:, 0:1])
    fgm = torch.zeros_like(target[:, 0:1])
    B = target.shape[0]
    ids_all = panoptics.reshape(B, *target.shape[2:])
    for i in range(B):
        for inst in torch.unique(ids_all[i]).tolist():
            if inst == 0:
                continue
            m = ids_all[i] == inst
            d_i = pds[i][:, m].t()
            g_i = gts[i][:, m].t()
            a_i = fn(d_i, g_i, **cfg)
            k = min(int(cfg["k"]), len(a_i))
            val, idx = a_i.topk(k)
            like = torch.zeros_like(a_i).scatter(0, idx, val)
            aff[i, 0][m] = like.type_as(aff)
            fgm[i, 0][m] = like.bool().type_as(aff)
    bgm = torch.logical_and(fgm.logical_not(), mask)
    return aff * onehot, fgm, bgm, onehot.any(dim=1, keepdim=True)
