"""CPU restatement of the reference's rasterize -> decode -> NMS path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Citations are relative to
/root/reference.  numpy for the rasterizer (the reference is numpy + numba there),
torch-on-CPU for decode / NMS control flow (the reference is torch there, so the
CPU baseline keeps the reference's threading behaviour), C (oracle/csrc/oracle.c)
for the serial z-buffer loop and the third-party IoU / NMS kernels.

Decisions the reference leaves open, fixed here (and mirrored by the CUDA path):
  * total order inside a (sweep, class): score descending, then candidate index
    ascending (``topk`` tie order is unspecified upstream);
  * hard NMS suppresses on ``iou > thr`` (detectron2's CUDA comparison; its CPU
    kernel uses ``>=``), thr = float32(iou_threshold) because the reference passes
    ``torch.as_tensor(iou_threshold)`` (math/ops/nms.py:44);
  * weighted NMS: see orc_wnms in oracle.c (TorchEx is absent: parity unpinned).
"""
from __future__ import annotations

import ctypes
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .build import lib

__all__ = [
    "ROW_MAPPING_64", "cart_to_sph", "build_range_view_coordinates",
    "build_range_view_coordinates_converter", "z_buffer", "build_range_view",
    "decode_range_view", "sample_by_range", "bchw_to_bkc", "yaw_to_quat",
    "rot_iou_pairs", "mmcv_iou_pairs", "nms_rotated", "iou_bev_pairs", "weighted_nms",
    "hard_multiclass_nms", "weighted_multiclass_nms", "batched_multiclass_nms",
    "range_decoder_decode", "iou_3d_axis_aligned", "subsample_range_view",
]

# src/torchbox3d/prototype/loader.py:62-129 (= datasets/argoverse/constants.py:560-627)
ROW_MAPPING_64 = np.array(
    [56, 22, 42, 28, 61, 30, 49, 36, 40, 32, 38, 45, 34, 26, 53, 59, 8, 1, 16, 20, 12, 5,
     11, 15, 17, 9, 24, 6, 13, 3, 19, 0, 7, 41, 21, 35, 2, 33, 14, 27, 23, 31, 25, 18, 29,
     37, 10, 4, 55, 62, 47, 43, 51, 58, 52, 48, 46, 54, 39, 57, 50, 60, 44, 63])


def _p(a: np.ndarray) -> ctypes.c_void_p:
    return ctypes.c_void_p(a.ctypes.data)


# --------------------------------------------------------------------------- #
# rasterize                                                                    #
# --------------------------------------------------------------------------- #
def cart_to_sph(cart: np.ndarray) -> np.ndarray:
    """math/numpy/conversions.py:46-73 -> (N,3) [azimuth, inclination, radius]."""
    x, y, z = cart[..., 0], cart[..., 1], cart[..., 2]
    hxy = np.hypot(x, y)
    out = np.zeros_like(cart)
    out[..., 0] = np.arctan2(y, x)
    out[..., 1] = np.arctan2(z, hxy)
    out[..., 2] = np.hypot(hxy, z)
    return out


def build_range_view_coordinates(cart, sph, laser_numbers, laser_mapping,
                                 n_inclination_bins: int = 64, n_azimuth_bins: int = 1800):
    """math/numpy/conversions.py:9-43.  NOTE: mutates ``sph[..., 0]`` in place, as the
    reference does (:30-34); col = round_half_even(W - az' - 1), clipped to [0, W-1]."""
    az = sph[..., 0]
    az += math.pi
    az *= n_azimuth_bins / math.tau
    col = np.clip((n_azimuth_bins - az - 1).round(), 0, n_azimuth_bins - 1)
    row = n_inclination_bins - laser_mapping[laser_numbers] - 1
    hybrid = np.zeros_like(cart)
    hybrid[:, 0], hybrid[:, 1], hybrid[:, 2] = row, col, sph[..., 2]
    return hybrid


def build_range_view_coordinates_converter(cart, sph, laser_numbers, laser_mapping,
                                           n_inclination_bins: int, n_azimuth_bins: int,
                                           build_uniform_inclination: bool = False):
    """converters/av2/utils.py:108-153: col = W - round(az'); optional uniform rows."""
    az, inc = sph[..., 0], sph[..., 1]
    az += math.pi
    az *= n_azimuth_bins / math.tau
    col = n_azimuth_bins - np.round(az)
    if build_uniform_inclination:
        fov = np.abs(np.array([-10.0 / 180.0 * math.pi, 10 / 180.0 * math.pi]))
        row = 1.0 - (inc + fov[0]) / (fov[0] + fov[1])
        row = (row * n_inclination_bins).round().clip(0, n_inclination_bins - 1)
    else:
        row = n_inclination_bins - laser_mapping[laser_numbers] - 1
    col = np.clip(col, 0, n_azimuth_bins - 1)
    hybrid = np.zeros_like(cart)
    hybrid[:, 0], hybrid[:, 1], hybrid[:, 2] = row, col, sph[..., 2]
    return hybrid


def z_buffer(indices, distances, features, height: int, width: int,
             min_distance: float = 1.0, return_winner: bool = False):
    """math/numpy/conversions.py:106-128 (serial loop in C: oracle.c orc_zbuffer)."""
    rows = np.ascontiguousarray(indices[0], dtype=np.int64)
    cols = np.ascontiguousarray(indices[1], dtype=np.int64)
    dist = np.ascontiguousarray(distances)
    feat = np.ascontiguousarray(features)
    assert dist.dtype in (np.float64, np.float32) and feat.dtype in (np.float64, np.float32)
    C, N = feat.shape
    image = np.empty((C, height * width), dtype=np.float32)
    winner = np.empty(height * width, dtype=np.int32)
    lib().orc_zbuffer(_p(rows), _p(cols), _p(dist), int(dist.dtype == np.float64), _p(feat),
                      int(feat.dtype == np.float64), C, N, height, width, float(min_distance),
                      _p(image), _p(winner))
    image = image.reshape(C, height, width)
    return (image, winner.reshape(height, width)) if return_winner else image


def build_range_view(xyz, intensity, laser_number, laser_mapping, lidar_offset,
                     num_lasers: int = 64, width: int = 1800, n_azimuth_bins: int = 1800,
                     return_winner: bool = False):
    """math/range_view.py:14-44 over plain arrays (the reference takes a polars frame).

    The reference never forwards ``width`` to build_range_view_coordinates (:34-40), so
    the column is always computed for 1800 bins; ``n_azimuth_bins`` exposes that.
    Returns (7,H,W) f32 [az, inc, r, x, y, z, intensity]; winner indices refer to the
    ORIGINAL (unfiltered) point order."""
    keep = np.nonzero(laser_number < num_lasers)[0]                       # :23-26
    xyz_k = xyz[keep]
    cart = xyz_k - np.asarray(lidar_offset)                                # :29
    sph = cart_to_sph(cart)                                                # :30
    feats = np.concatenate([sph, xyz_k, intensity[keep].reshape(-1, 1)], axis=1).T  # :33 (copy)
    hybrid = build_range_view_coordinates(cart, sph, laser_number[keep].astype(np.int64),
                                          np.asarray(laser_mapping), num_lasers, n_azimuth_bins)
    indices = np.ascontiguousarray(hybrid[:, :2].T.astype(int))            # :41
    out = z_buffer(indices, hybrid[:, 2], feats, num_lasers, width, return_winner=return_winner)
    if return_winner:
        img, win = out
        if len(keep):
            win = np.where(win >= 0, keep[np.maximum(win, 0)], -1).astype(np.int32)
        return img, win
    return out


def subsample_range_view(range_view, mask, cart, dataset_name: str, x_stride: int, mode: str):
    """prototype/loader.py:792-815."""
    pad = {"waymo": [19, 19] if x_stride == 4 else [3, 3],
           "av2": [28, 28] if x_stride == 4 else [4, 4]}[dataset_name]
    f = torch.nn.functional.pad
    range_view = range_view * mask
    return (f(range_view, pad, mode=mode)[:, :, ::x_stride], f(mask, pad, mode=mode)[:, :, ::x_stride],
            f(cart, pad, mode=mode)[:, :, ::x_stride])


# --------------------------------------------------------------------------- #
# decode                                                                       #
# --------------------------------------------------------------------------- #
def decode_range_view(regressands: torch.Tensor, cart: torch.Tensor,
                      enable_azimuth_invariant_targets: bool) -> torch.Tensor:
    """math/ops/coding.py:110-144 (+ egovehicle_from_azimuth :79-107): f64 inside."""
    dt = regressands.dtype
    r = regressands.double()
    c = cart.double()
    ox, oy, oz = r[:, 0:1], r[:, 1:2], r[:, 2:3]
    lwh = r[:, 3:6].exp()
    yaw = torch.atan2(r[:, 6:7], r[:, 7:8])
    if enable_azimuth_invariant_targets:
        phi = torch.atan2(c[:, 1:2], c[:, 0:1])
        s, co = phi.sin(), phi.cos()
        ox, oy = co * ox - s * oy, s * ox + co * oy
        yaw = yaw + phi
    ctr = c + torch.cat([ox, oy, oz], dim=1)
    return torch.cat([ctr, lwh, yaw], dim=1).to(dt)


def bchw_to_bkc(x: torch.Tensor) -> torch.Tensor:
    """math/conversions.py:174-186."""
    return x.permute(0, 2, 3, 1).reshape(x.shape[0], -1, x.shape[1])


def sample_by_range(scores, categories, cuboids, cart, lower_bounds, upper_bounds, subsampling_rates):
    """nn/decoders/range_decoder.py:127-156.  scores/categories (B,1,H,W), cuboids (B,7,H,W)."""
    d = cart.norm(dim=1, keepdim=True)                                              # :140
    ss, cs, bs = [], [], []
    for lo, hi, rate in zip(lower_bounds, upper_bounds, subsampling_rates):
        part = torch.logical_and(d > lo, d <= hi)                                   # :143
        ss.append((scores * part)[:, :, :, ::rate].flatten(2))                      # only scores are masked
        cs.append(categories[:, :, :, ::rate].flatten(2))
        bs.append(cuboids[:, :, :, ::rate].flatten(2))
    return torch.cat(ss, -1).squeeze(1), torch.cat(cs, -1).squeeze(1), torch.cat(bs, -1).transpose(2, 1)


def yaw_to_quat(yaw: torch.Tensor) -> torch.Tensor:
    """math/linalg/lie/SO3.py:122-134 via kornia quaternion_from_euler(0,0,yaw) [parity unpinned]:
    (qw,qx,qy,qz) = (cos(yaw/2), 0, 0, sin(yaw/2))."""
    h = yaw * 0.5
    z = torch.zeros_like(h)
    return torch.cat([h.cos(), z, z, h.sin()], dim=-1)


# --------------------------------------------------------------------------- #
# IoU / NMS kernels (C)                                                        #
# --------------------------------------------------------------------------- #
def rot_iou_pairs(a: np.ndarray, b: np.ndarray, angle_scale: float) -> np.ndarray:
    """Aligned rotated IoU of (n,5) f32 (xc,yc,w,h,angle) boxes, detectron2's routine and rotation direction (w-axis
    along (cos, -sin)); angle_scale = 0.01745329251 for degrees in (detectron2's own unit), 1.0 for radians."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty(a.shape[0], dtype=np.float32)
    lib().orc_rot_iou_aligned(_p(a), _p(b), a.shape[0], float(angle_scale), _p(out))
    return out


def mmcv_iou_pairs(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """mmcv.ops.box_iou_rotated(a, b, aligned=True) with its default clockwise=True: angles in radians, w-axis along
    (cos, +sin) -- the same routine as detectron2's with the other rotation direction (pinned by mmcv's published
    unit-test vector, tests/test_oracle_iou.py::test_mmcv_published_vectors)."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    out = np.empty(a.shape[0], dtype=np.float32)
    lib().orc_rot_iou_aligned_mmcv(_p(a), _p(b), a.shape[0], _p(out))
    return out


def _order_desc(scores: np.ndarray) -> np.ndarray:
    # score descending, index ascending on ties
    return np.argsort(-scores.astype(np.float64), kind="stable").astype(np.int64)


def nms_rotated(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold, return_evals: bool = False,
                max_keep: int = -1):
    """Stand-in for detectron2.layers.nms.nms_rotated (call site math/ops/nms.py:41-45):
    (N,5) f32 (xc,yc,w,h,angle_deg), (N,) f32 -> kept ORIGINAL indices in score order."""
    b = np.ascontiguousarray(boxes.detach().cpu().numpy(), dtype=np.float32)
    s = np.ascontiguousarray(scores.detach().cpu().numpy(), dtype=np.float32)
    thr = float(iou_threshold)   # a 0-dim f32 tensor -> double(f32 value), as upstream
    order = _order_desc(s)
    keep = np.empty(max(len(s), 1), dtype=np.int64)
    evals = ctypes.c_int64(0)
    n = lib().orc_nms_rotated(_p(b), _p(order), len(s), thr, 0.01745329251, _p(keep), ctypes.byref(evals), max_keep)
    out = torch.from_numpy(keep[:n].copy())
    return (out, evals.value) if return_evals else out


def iou_bev_pairs(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    L = lib()
    return np.array([L.orc_iou_bev(_p(a[i]), _p(b[i])) for i in range(a.shape[0])], dtype=np.float32)


def weighted_nms(boxes: torch.Tensor, data2merge: torch.Tensor, scores: torch.Tensor,
                 nms_threshold: float, merge_thresh: float, max_out: int = -1):
    """math/ops/nms.py:126-177 with wnms_gpu replaced by oracle.c orc_wnms."""
    s = scores.detach().cpu().numpy().astype(np.float32)
    order = _order_desc(s)
    b = np.ascontiguousarray(boxes.detach().cpu().numpy()[order], dtype=np.float32)
    d = data2merge.detach().cpu().numpy()[order].astype(np.float32)
    ds = np.ascontiguousarray(np.concatenate([d, s[order][:, None]], axis=1), dtype=np.float32)
    n, D = ds.shape
    output = np.zeros_like(ds)
    keep = np.zeros(max(n, 1), dtype=np.int64)
    count = np.zeros(max(n, 1), dtype=np.int64)
    m = lib().orc_wnms(_p(b), _p(ds), n, D, np.float32(nms_threshold), np.float32(merge_thresh),
                       _p(output), _p(keep), _p(count), max_out)
    return (torch.from_numpy(order[keep[:m]]), torch.from_numpy(output[:m].copy()),
            torch.from_numpy(count[:m].copy()))


# --------------------------------------------------------------------------- #
# NMS control flow                                                             #
# --------------------------------------------------------------------------- #
def _topk_stable(scores: torch.Tensor, k: int):
    v, i = torch.sort(scores, descending=True, stable=True)
    return v[:k], i[:k]


def hard_multiclass_nms(cuboids_i, scores_i, categories_i, iou_threshold, num_pre_nms, num_post_nms):
    """math/ops/nms.py:11-61."""
    outs: List[Tuple[torch.Tensor, torch.Tensor, torch.Tensor]] = []
    for j in torch.unique(categories_i):
        m = categories_i == j
        sc, cu = scores_i[m], cuboids_i[m]
        sc, rank = _topk_stable(sc, min(len(sc), num_pre_nms))                       # :29-32
        cu = cu[rank]
        inp = cu[:, [0, 1, 3, 4, 6]].contiguous().float()
        inp[:, -1] = -inp[:, -1].rad2deg()                                           # :40 (f32)
        # :41-45; max_keep: only the first num_post_nms kept boxes survive :53-56, so stopping there is exact
        keep = nms_rotated(inp, sc.float(), torch.as_tensor(iou_threshold), max_keep=num_post_nms)
        cu, sc = cu[keep], sc[keep].flatten()
        sc, rank = _topk_stable(sc, min(len(cu), num_post_nms))                      # :53-56
        outs.append((cu[rank], sc, torch.full_like(sc, fill_value=float(j))))
    return tuple(torch.cat(x) for x in zip(*outs))


def weighted_multiclass_nms(cuboids_i, scores_i, categories_i, iou_threshold, num_pre_nms, num_post_nms):
    """math/ops/nms.py:64-123."""
    outs = []
    for j in categories_i.unique():
        m = categories_i == j
        sc, cu = scores_i[m], cuboids_i[m]
        sc, rank = _topk_stable(sc, min(len(sc), num_pre_nms))
        cu = cu[rank]
        bx = cu[..., [0, 1, 3, 4, 6]].contiguous()
        inp = torch.cat([bx[:, :2] - bx[:, 2:4] / 2, bx[:, :2] + bx[:, 2:4] / 2, bx[:, -1:]], dim=-1)
        d2m = torch.cat([cu[:, :-1], cu[:, -1:].sin(), cu[:, -1:].cos()], dim=1)
        _, merged, _ = weighted_nms(inp, d2m, sc, iou_threshold, 0.5, max_out=num_post_nms)
        box6, sn, cs, sc = merged.split([6, 1, 1, 1], dim=1)
        cu = torch.cat([box6, torch.atan2(sn, cs)], dim=1)
        sc = sc.flatten()
        sc, rank = _topk_stable(sc, min(len(cu), num_post_nms))
        outs.append((cu[rank], sc, torch.full_like(sc, fill_value=float(j))))
    return tuple(torch.cat(x) for x in zip(*outs))


def batched_multiclass_nms(cuboids, scores, categories, num_pre_nms, num_post_nms, iou_threshold,
                           min_confidence, nms_mode):
    """math/ops/nms.py:181-266."""
    mode = nms_mode.upper()
    if mode not in ("HARD", "WEIGHTED"):
        raise NotImplementedError(f"NMS Mode: {mode} is not implemented.")
    fn = hard_multiclass_nms if mode == "HARD" else weighted_multiclass_nms
    cu_l, sc_l, ca_l, bi_l = [], [], [], []
    for i in range(cuboids.shape[0]):
        m = scores[i] >= min_confidence                                              # :212
        if int(m.sum()) == 0:
            continue
        cu, sc, ca = fn(cuboids[i, m], scores[i, m], categories[i, m], iou_threshold, num_pre_nms, num_post_nms)
        cu_l.append(cu); sc_l.append(sc); ca_l.append(ca); bi_l.append(torch.full_like(sc, fill_value=i))
    if not cu_l:
        return (cuboids.new_empty((0, cuboids.shape[-1])), scores.new_empty((0, 1)),
                categories.new_empty((0, 1)), categories.new_empty((0, 1)))
    return torch.cat(cu_l), torch.cat(sc_l), torch.cat(ca_l), torch.cat(bi_l)


def range_decoder_decode(multiscale_outputs: Dict, post_processing_config: Dict, task_config: Dict,
                         enable_azimuth_invariant_targets: bool, enable_sample_by_range: bool,
                         lower_bounds: Sequence[float], upper_bounds: Sequence[float],
                         subsampling_rates: Sequence[int], use_nms: bool = True,
                         return_candidates: bool = False):
    """nn/decoders/range_decoder.py:29-124."""
    sc_l, cu_l, ca_l = [], [], []
    for _stride, ms in multiscale_outputs.items():
        cart, mask = ms["cart"], ms["mask"]
        task_offset = 0
        for task_id, group in task_config.items():
            out = ms[task_id]
            scores = out["logits"].sigmoid() * mask                                   # :49-50
            scores, cats = scores.max(dim=1, keepdim=True)                            # :52-54
            cub = decode_range_view(out["regressands"], cart, enable_azimuth_invariant_targets)
            if enable_sample_by_range:
                scores, cats, cub = sample_by_range(scores, cats, cub, cart, tuple(lower_bounds),
                                                    tuple(upper_bounds), tuple(subsampling_rates))
            else:
                scores, cub, cats = bchw_to_bkc(scores).squeeze(-1), bchw_to_bkc(cub), bchw_to_bkc(cats).squeeze(-1)
            cats = cats + task_offset
            task_offset += len(group)
            sc_l.append(scores); cu_l.append(cub); ca_l.append(cats)
    params, scores, cats = torch.cat(cu_l, 1), torch.cat(sc_l, 1), torch.cat(ca_l, 1)
    if return_candidates:
        return params, scores, cats
    if use_nms:
        params, scores, cats, bidx = batched_multiclass_nms(
            params, scores, cats, post_processing_config["num_pre_nms"], post_processing_config["num_post_nms"],
            post_processing_config["nms_threshold"], post_processing_config["min_confidence"],
            post_processing_config["nms_mode"])
    else:
        B, N, _ = params.shape
        bidx = torch.arange(B).repeat_interleave(N)
        params, scores, cats = params.flatten(0, 1), scores.flatten(0, 1), cats.flatten(0, 1)
        t = scores >= post_processing_config["min_confidence"]
        params, scores, cats, bidx = params[t], scores[t], cats[t], bidx[t]
    params = torch.cat([params[:, :-1], yaw_to_quat(params[:, -1:])], dim=-1)
    return params, scores, cats, bidx


# --------------------------------------------------------------------------- #
# aligned 3D IoU                                                               #
# --------------------------------------------------------------------------- #
def iou_3d_axis_aligned(a: torch.Tensor, b: torch.Tensor):
    """math/ops/iou.py:11-47 with mmcv box_iou_rotated(aligned=True) -> orc_rot_iou_aligned_mmcv."""
    idx = [0, 1, 3, 4, 6]
    iou_bev = torch.from_numpy(mmcv_iou_pairs(a[:, idx].float().numpy(), b[:, idx].float().numpy()))
    iou_bev = iou_bev.clamp(0.0, 1.0).nan_to_num(nan=0.0)
    area_a, area_b = a[:, [3, 4]].prod(-1), b[:, [3, 4]].prod(-1)
    ov_bev = iou_bev * (area_a + area_b) / (1.0 + iou_bev)
    top = torch.min(a[:, 2] + a[:, 5] / 2.0, b[:, 2] + b[:, 5] / 2.0)
    btm = torch.max(a[:, 2] - a[:, 5] / 2.0, b[:, 2] - b[:, 5] / 2.0)
    ov3 = ov_bev * torch.clamp(top - btm, min=0)
    iou3 = ov3 / torch.clamp(a[:, 3:6].prod(-1) + b[:, 3:6].prod(-1) - ov3, min=1e-8)
    iou3 = iou3.nan_to_num(nan=0.0)
    if not iou3.isfinite().all():
        raise RuntimeError("Invalid IoUs.")
    return iou3, iou_bev
