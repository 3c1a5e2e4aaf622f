"""Host-side mirror of ``torchbox3d/math/ops/iou.py``."""
from __future__ import annotations

from typing import Tuple

import torch
from torch import Tensor

from ... import _native as N
from ..._util import ptr, require_cuda, stream_ptr

__all__ = ["iou_3d_axis_aligned", "iou"]


def iou_3d_axis_aligned(cuboids_a: Tensor, cuboids_b: Tensor) -> Tuple[Tensor, Tensor]:
    """Drop-in for math/ops/iou.py:11-47: aligned pairs of (N,7) cuboids -> (iou_3d (N,), iou_bev (N,));
    raises RuntimeError("Invalid IoUs.") on a non-finite result like upstream (:40-46)."""
    dev = require_cuda(cuboids_a, cuboids_b)
    n = cuboids_a.shape[0]
    if cuboids_b.shape[0] != n:
        raise ValueError("aligned IoU needs the same number of cuboids on both sides")
    a = cuboids_a[:, :7].float().contiguous()
    b = cuboids_b[:, :7].float().contiguous()
    iou3d = torch.empty((n,), dtype=torch.float32, device=dev)
    bev = torch.empty((n,), dtype=torch.float32, device=dev)
    status = torch.zeros((1,), dtype=torch.int32, device=dev)
    N.check(N.lib().rv3d_iou3d_aligned(ptr(a), ptr(b), n, ptr(iou3d), ptr(bev), ptr(status), stream_ptr(dev)),
            "rv3d_iou3d_aligned")
    if int(status.item()) != 0:
        raise RuntimeError("Invalid IoUs.")
    return iou3d, bev


def iou(src_dims_m: Tensor, target_dims_m: Tensor) -> Tensor:
    """math/ops/iou.py:50-55 (pure index arithmetic on two small tensors; not on the hot path)."""
    inter = torch.minimum(src_dims_m, target_dims_m).prod(axis=1, keepdim=True)
    union = torch.maximum(src_dims_m, target_dims_m).prod(axis=1, keepdim=True)
    return torch.divide(inter, union)
