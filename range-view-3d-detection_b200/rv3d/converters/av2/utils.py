"""Mirror of ``converters/av2/utils.py`` (the production exporter's sweep preparation) over
arrays / tensors instead of polars frames: same function names, same arithmetic, computed on the GPU in fp64.

    unmotion_compensate    utils.py:229-295
    sensor_from_egovehicle utils.py:43-57   (SE3(R, t).inverse().transform_point_cloud)
    correct_laser_numbers  utils.py:211-226
    build_range_view       utils.py:32-105  (the three above feed cart_to_sph / build_range_view_coordinates /
                                             z_buffer, which already live in rv3d.math.numpy.conversions)
"""
from __future__ import annotations

from typing import Collection, Optional, Tuple

import numpy as np
import torch

from ... import _native as N
from ..._util import ptr, stream_ptr
from ...constants import LASER_MAPPING, ROW_MAPPING_32, ROW_MAPPING_64
from ...math.numpy.conversions import _back, _dev, _to_dev, build_range_view_coordinates, cart_to_sph, z_buffer

__all__ = ["unmotion_compensate", "sensor_from_egovehicle", "correct_laser_numbers", "build_range_view"]


def _host_f64(x, shape) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(shape))


def unmotion_compensate(xyz, offset_ns, timestamp_ns: int, pose_timestamps_ns, pose_quat_xyzw, pose_translation,
                        device=None) -> Tuple:
    """utils.py:229-295 -> (xyz_p (N',3) float64, keep (N,) bool).

    ``xyz`` (N,3), ``offset_ns`` (N,) are the sweep's columns; the three pose arrays are the log's
    ``city_SE3_egovehicle`` table sorted by time (``timestamp_ns``, the ``(qx,qy,qz,qw)`` columns in that order,
    ``(tx_m,ty_m,tz_m)``).  Rows whose time is not strictly inside the table are dropped like the reference's
    ``filter`` (``keep`` marks the survivors so the caller can filter its other columns).  Raises ``ValueError``
    when no pose carries the sweep's own ``timestamp_ns`` (the reference fails on the empty selection)."""
    dev = _dev(device)
    pts = _to_dev(xyz, torch.float64, dev).reshape(-1, 3)
    off = _to_dev(offset_ns, torch.int64, dev).reshape(-1)
    ts_host = (pose_timestamps_ns.detach().cpu().numpy() if isinstance(pose_timestamps_ns, torch.Tensor)
               else np.asarray(pose_timestamps_ns)).astype(np.int64)
    hit = np.nonzero(ts_host == int(timestamp_ns))[0]
    if hit.size == 0:
        raise ValueError(f"no pose at the sweep's timestamp {timestamp_ns}")
    quat_host = _host_f64(pose_quat_xyzw, (-1, 4))
    trans_host = _host_f64(pose_translation, (-1, 3))
    if ts_host.shape[0] != quat_host.shape[0] or ts_host.shape[0] != trans_host.shape[0] or ts_host.shape[0] < 2:
        raise ValueError("pose table: need >= 2 rows and equally long timestamp / quaternion / translation arrays")
    if pts.shape[0] != off.shape[0]:
        raise ValueError("xyz and offset_ns disagree in length")
    ts = torch.as_tensor(ts_host, device=dev)
    quat = torch.as_tensor(quat_host, device=dev)
    trans = torch.as_tensor(trans_host, device=dev)
    tq = np.ascontiguousarray(quat_host[hit[0]])
    tt = np.ascontiguousarray(trans_host[hit[0]])
    out = torch.empty_like(pts)
    valid = torch.empty((pts.shape[0],), dtype=torch.uint8, device=dev)
    N.check(N.lib().rv3d_unmotion_compensate(ptr(pts), ptr(off), pts.shape[0], int(timestamp_ns), ptr(ts), ptr(quat),
                                             ptr(trans), ts.shape[0], tq.ctypes.data, tt.ctypes.data, ptr(out),
                                             ptr(valid), stream_ptr(dev)), "rv3d_unmotion_compensate")
    keep = valid.bool()
    return _back(out[keep], xyz), _back(keep, xyz)


def sensor_from_egovehicle(xyz, rotation, translation, device=None):
    """utils.py:43-57: ``SE3(rotation, translation).inverse().transform_point_cloud(xyz)`` with the (3,3) rotation of
    ``egovehicle_SE3_sensor`` (``Rotation.from_quat(qx,qy,qz,qw).as_matrix()``) and its translation."""
    dev = _dev(device)
    pts = _to_dev(xyz, torch.float64, dev).reshape(-1, 3)
    rot = _host_f64(rotation, (9,))
    tr = _host_f64(translation, (3,))
    out = torch.empty_like(pts)
    N.check(N.lib().rv3d_transform_points(ptr(pts), pts.shape[0], rot.ctypes.data, tr.ctypes.data, 1, ptr(out),
                                          stream_ptr(dev)), "rv3d_transform_points")
    return _back(out, xyz)


def correct_laser_numbers(laser_numbers, log_id: str, height: int, log_ids: Optional[Collection[str]] = None,
                          device=None):
    """utils.py:211-226 -> row-mapped laser numbers (int64).  ``log_ids`` is the reference's ``LOG_IDS`` table
    (``datasets/argoverse/constants.py:269``: the logs recorded with the other beam ordering); it is dataset
    metadata, not arithmetic, and stays with the caller.  Out-of-table laser numbers raise ``IndexError`` like numpy."""
    dev = _dev(device)
    las = _to_dev(laser_numbers, torch.int64, dev).reshape(-1)
    remap = log_ids is not None and log_id in log_ids
    mapping = torch.as_tensor(LASER_MAPPING.astype(np.int64), device=dev) if remap else None
    rows = torch.as_tensor((ROW_MAPPING_32 if height == 32 else ROW_MAPPING_64).astype(np.int64), device=dev)
    out = torch.empty_like(las)
    bad = torch.zeros(1, dtype=torch.int32, device=dev)
    N.check(N.lib().rv3d_correct_laser_numbers(ptr(las), las.shape[0], ptr(mapping), ptr(rows), rows.shape[0], ptr(out),
                                               ptr(bad), stream_ptr(dev)), "rv3d_correct_laser_numbers")
    if int(bad.item()):
        raise IndexError("laser number outside the mapping tables")
    return _back(out, laser_numbers)


def build_range_view(cart, features, laser_number, offset_ns, rotation, translation, height: int, width: int,
                     build_uniform_inclination: bool = False, device=None, return_winner: bool = False):
    """utils.py:32-105 -> (8, H, W) float32 ``[x, y, z, intensity, laser_number, is_within_roi, timedelta_ns, range]``
    (the reference wraps the same array, flattened, in a polars frame with ``RANGE_VIEW_SCHEMA``).

    ``cart`` (N,3): the ``x_p, y_p, z_p`` columns; ``features`` (N,6): the ``FEATURE_COLUMN_NAMES`` columns;
    ``rotation`` (3,3) / ``translation`` (3,): ``egovehicle_SE3_sensor`` of the lidar."""
    dev = _dev(device)
    cart_lidar = sensor_from_egovehicle(_to_dev(cart, torch.float64, dev), rotation, translation, dev)   # :43-57
    las = _to_dev(laser_number, torch.int64, dev).reshape(-1)                                            # :59-61
    t_ns = _to_dev(offset_ns, torch.float64, dev).reshape(-1)                                            # :63
    sph = cart_to_sph(cart_lidar, dev)                                                                   # :68
    mapping = torch.arange(height, dtype=torch.int64, device=dev)                                        # :67
    hybrid = build_range_view_coordinates(cart_lidar, sph, las, mapping, height, width, dev,             # :69-77
                                          "converter_uniform" if build_uniform_inclination else "converter")
    indices = hybrid[:, :2].to(torch.int64).T.contiguous()                                               # :79
    distances = hybrid[:, 2].contiguous()                                                                # :80
    feats = torch.cat((_to_dev(features, torch.float64, dev), t_ns[:, None], distances[:, None]), dim=-1).T.contiguous()  # :83-85
    res = z_buffer(indices, distances, feats, height, width, device=dev, return_winner=return_winner)    # :90
    if return_winner:
        return _back(res[0], cart), _back(res[1], cart)
    return _back(res, cart)
