"""Host-side mirror of the range-view post-processing in ``torchbox3d/prototype/loader.py``:
the step between the rasterizer's output and the backbone's input, as one CUDA pass."""
from __future__ import annotations

from typing import Sequence, Tuple

import torch
from torch import Tensor

from .. import _native as N
from .._util import ptr, require_cuda, stream_ptr

__all__ = ["subsample_range_view", "range_view_inputs", "rasterize_inputs", "intersection_test", "IMAGE_CHANNELS"]

# channel order of rv3d.math.range_view images (math/range_view.py:33)
IMAGE_CHANNELS = ("azimuth", "inclination", "range", "x", "y", "z", "intensity")


def _pad_for(dataset_name: str, x_stride: int) -> int:
    """prototype/loader.py:800-809."""
    if dataset_name == "waymo":
        return 19 if x_stride == 4 else 3
    if dataset_name == "av2":
        return 28 if x_stride == 4 else 4
    raise ValueError(f"unknown dataset {dataset_name!r}")


def _inputs_params(B: int, H: int, W: int, feature_column_names: Sequence[str], dataset_name: str, x_stride: int, mode: str):
    if mode not in ("circular", "constant"):
        raise NotImplementedError(f"padding mode {mode!r}")
    p = N.InputsParams()
    p.batch, p.height, p.width = B, H, W
    p.x_stride, p.pad = int(x_stride), _pad_for(dataset_name, x_stride)
    p.pad_mode = 0 if mode == "circular" else 1
    p.n_features = len(feature_column_names)
    for f, name in enumerate(feature_column_names):
        p.feature_channel[f] = IMAGE_CHANNELS.index(name)
    p.tanh_channel = IMAGE_CHANNELS.index("intensity") if dataset_name == "waymo" else -1
    return p, (W + 2 * p.pad + p.x_stride - 1) // p.x_stride


def rasterize_inputs(points: Tensor, laser: Tensor, n_points: Tensor, laser_mapping: Tensor, lidar_offset: Sequence[float],
                     height: int = 64, width: int = 1800, feature_column_names: Sequence[str] = ("intensity", "range", "x", "y", "z"),
                     dataset_name: str = "av2", x_stride: int = 1, mode: str = "circular", n_azimuth_bins: int = None,
                     num_lasers: int = None, col_mode: str = "library", min_distance: float = 1.0,
                     workspace: Tensor = None) -> Tuple[Tensor, Tensor, Tensor]:
    """Raw sweeps -> (features (B,F,H,Wo), mask (B,1,H,Wo) bool, cart (B,3,H,Wo)) in one scatter + one resolve pass:
    ``rv3d.math.range_view.rasterize_sweeps`` followed by ``range_view_inputs`` (math/range_view.py:14-44 +
    prototype/loader.py:623-650, 792-815) without materialising the 7-plane range image in between.  Arguments as in
    those two functions; results are identical to calling them one after the other."""
    from ..math.range_view import raster_params
    dev = require_cuda(points, laser, n_points, laser_mapping)
    if points.dtype != torch.float32 or points.dim() != 3 or points.shape[-1] != 4:
        raise ValueError("points must be (B,Nmax,4) float32")
    if laser.dtype != torch.uint8 or n_points.dtype != torch.int32 or laser_mapping.dtype != torch.int32:
        raise ValueError("laser must be uint8; n_points and laser_mapping int32")
    points, laser = points.contiguous(), laser.contiguous()
    B = points.shape[0]
    rp, need = raster_params(B, points.shape[1], height, width, n_azimuth_bins, num_lasers, laser_mapping.numel(), col_mode,
                             lidar_offset, min_distance)
    ip, wo = _inputs_params(B, height, width, feature_column_names, dataset_name, x_stride, mode)
    if workspace is None or workspace.numel() * workspace.element_size() < need:
        from .._util import scratch
        workspace = scratch(need, dev)
    features = torch.empty((B, ip.n_features, height, wo), dtype=torch.float32, device=dev)
    cart = torch.empty((B, 3, height, wo), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 1, height, wo), dtype=torch.uint8, device=dev)
    N.check(N.lib().rv3d_rasterize_inputs(rp, ip, ptr(points), ptr(laser), ptr(n_points), ptr(laser_mapping), ptr(features),
                                          ptr(cart), ptr(mask), ptr(workspace), workspace.numel() * workspace.element_size(),
                                          stream_ptr(dev)), "rv3d_rasterize_inputs")
    return features, mask.view(torch.bool), cart


def range_view_inputs(image: Tensor, feature_column_names: Sequence[str] = ("intensity", "range", "x", "y", "z"),
                      dataset_name: str = "av2", x_stride: int = 1, mode: str = "circular"
                      ) -> Tuple[Tensor, Tensor, Tensor]:
    """Rasterized image (B,7,H,W) -> (features (B,F,H,Wo), mask (B,1,H,Wo) bool, cart (B,3,H,Wo)): the
    loader's feature / cart / mask assembly (prototype/loader.py:623-650; Waymo ``tanh(intensity)`` :625-626;
    ``mask = range > 0`` :645-650) fused with ``subsample_range_view`` (:792-815)."""
    dev = require_cuda(image)
    if image.dim() != 4 or image.shape[1] != 7 or image.dtype != torch.float32:
        raise ValueError("image must be (B,7,H,W) float32 as produced by rasterize_sweeps")
    B, _, H, W = image.shape
    p, wo = _inputs_params(B, H, W, feature_column_names, dataset_name, x_stride, mode)
    features = torch.empty((B, p.n_features, H, wo), dtype=torch.float32, device=dev)
    cart = torch.empty((B, 3, H, wo), dtype=torch.float32, device=dev)
    mask = torch.empty((B, 1, H, wo), dtype=torch.uint8, device=dev)
    N.check(N.lib().rv3d_range_view_inputs(p, ptr(image.contiguous()), ptr(features), ptr(cart), ptr(mask),
                                           stream_ptr(dev)), "rv3d_range_view_inputs")
    return features, mask.view(torch.bool), cart


def subsample_range_view(range_view: Tensor, mask: Tensor, cart: Tensor, dataset_name: str, x_stride: int, mode: str
                         ) -> Tuple[Tensor, Tensor, Tensor]:
    """Drop-in for prototype/loader.py:792-815 on (C,H,W) CUDA tensors: ``range_view *= mask``, pad along W
    (``circular`` / ``constant``), keep every ``x_stride``-th column -> (range_view, mask, cart)."""
    dev = require_cuda(range_view, mask, cart)
    if mode not in ("circular", "constant"):
        raise NotImplementedError(f"padding mode {mode!r}")
    C, H, W = range_view.shape
    pad = _pad_for(dataset_name, x_stride)
    wo = (W + 2 * pad + x_stride - 1) // x_stride
    rv = range_view.float().contiguous()
    mk = (mask != 0).reshape(1, H, W).contiguous().view(torch.uint8)
    ct = cart.float().contiguous()
    o_rv = torch.empty((C, H, wo), dtype=torch.float32, device=dev)
    o_mk = torch.empty((1, H, wo), dtype=torch.uint8, device=dev)
    o_ct = torch.empty((3, H, wo), dtype=torch.float32, device=dev)
    N.check(N.lib().rv3d_subsample_range_view(ptr(rv), ptr(mk), ptr(ct), 1, C, H, W, int(x_stride), pad,
                                              0 if mode == "circular" else 1, ptr(o_rv), ptr(o_mk), ptr(o_ct),
                                              stream_ptr(dev)), "rv3d_subsample_range_view")
    out_mask = o_mk.view(torch.bool) if mask.dtype == torch.bool else o_mk.to(mask.dtype)
    return o_rv.to(range_view.dtype), out_mask, o_ct.to(cart.dtype)


def intersection_test(annotations_tch: Tensor, db_samples_tch: Tensor) -> Tensor:
    """prototype/loader.py:775-789 (`_intersection_test`, the GT-database collision test of the copy-paste
    augmentation) on the already converted tensors: rows are ``TCH_COLUMN_NAMES`` cuboids
    ``(x, y, z, l, w, h, ..., yaw)``; -> (N, M) rotated BEV IoU of ``[:, [0, 1, 3, 4, -1]]``."""
    from ..math.ops.assignment import box_iou_rotated
    cols = [0, 1, 3, 4, -1]
    return box_iou_rotated(annotations_tch.float()[:, cols], db_samples_tch.float()[:, cols])
