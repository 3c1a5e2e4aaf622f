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

__all__ = ["PoseTable", "unmotion_compensate", "sensor_from_egovehicle", "correct_laser_numbers", "build_range_view"]


def _host_f64(x, shape) -> np.ndarray:
    if isinstance(x, torch.Tensor):
        x = x.detach().cpu().numpy()
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64).reshape(shape))


class PoseTable:
    """A log's ``city_SE3_egovehicle`` table, resident on the device (one upload per log instead of one per sweep) with
    everything a Slerp step needs per pose pair precomputed (``rv3d_pose_intervals``; scipy's ``Slerp.__init__`` does
    the same once per construction, utils.py:251-256).  Pass it to ``unmotion_compensate`` in place of the three arrays."""

    def __init__(self, pose_timestamps_ns, pose_quat_xyzw, pose_translation, device=None):
        dev = _dev(device)
        ts_host = (pose_timestamps_ns.detach().cpu().numpy() if isinstance(pose_timestamps_ns, torch.Tensor)
                   else np.asarray(pose_timestamps_ns)).astype(np.int64)
        self.quat_host = _host_f64(pose_quat_xyzw, (-1, 4))
        self.trans_host = _host_f64(pose_translation, (-1, 3))
        m = ts_host.shape[0]
        if m != self.quat_host.shape[0] or m != self.trans_host.shape[0] or m < 2:
            raise ValueError("pose table: need >= 2 rows and equally long timestamp / quaternion / translation arrays")
        self.ts_host, self.device, self.n_poses = ts_host, dev, m
        self.first_ns, self.last_ns = int(ts_host[0]), int(ts_host[-1])
        if self.last_ns <= self.first_ns:
            raise ValueError("pose table: timestamps must be sorted and span a positive interval")
        self._sorted = bool(np.all(ts_host[1:] >= ts_host[:-1]))
        self.ts = torch.as_tensor(ts_host, device=dev)
        self.trans = torch.as_tensor(self.trans_host, device=dev)
        quat = torch.as_tensor(self.quat_host, device=dev)
        self.intervals = torch.empty((m - 1, 8), dtype=torch.float64, device=dev)
        N.check(N.lib().rv3d_pose_intervals(ptr(quat), m, ptr(self.intervals), stream_ptr(dev)), "rv3d_pose_intervals")

    def row_at(self, timestamp_ns: int) -> int:
        """Row of the pose stamped exactly ``timestamp_ns`` (utils.py:258-273; the first one, like the reference's
        ``filter(...)[0]``); ValueError when there is none."""
        t = int(timestamp_ns)
        if self._sorted:
            k = int(np.searchsorted(self.ts_host, t, side="left"))
            if k < self.n_poses and int(self.ts_host[k]) == t:
                return k
        else:
            hit = np.nonzero(self.ts_host == t)[0]
            if hit.size:
                return int(hit[0])
        raise ValueError(f"no pose at the sweep's timestamp {timestamp_ns}")


def unmotion_compensate(xyz, offset_ns, timestamp_ns: int, pose_timestamps_ns, pose_quat_xyzw=None, pose_translation=None,
                        device=None) -> Tuple:
    """utils.py:229-295 -> (xyz_p (N',3) float64, keep (N,) bool).

    ``xyz`` (N,3), ``offset_ns`` (N,) are the sweep's columns; the three pose arrays are the log's
    ``city_SE3_egovehicle`` table sorted by time (``timestamp_ns``, the ``(qx,qy,qz,qw)`` columns in that order,
    ``(tx_m,ty_m,tz_m)``) -- or ONE ``PoseTable`` built once per log in place of ``pose_timestamps_ns``.  Rows whose
    time is not strictly inside the table are dropped like the reference's ``filter`` (``keep`` marks the survivors so
    the caller can filter its other columns).  Raises ``ValueError`` when no pose carries the sweep's own
    ``timestamp_ns`` (the reference fails on the empty selection)."""
    table = pose_timestamps_ns if isinstance(pose_timestamps_ns, PoseTable) else None
    dev = table.ts.device if (table is not None and device is None) else _dev(device)
    if table is None:
        ts_host = (pose_timestamps_ns.detach().cpu().numpy() if isinstance(pose_timestamps_ns, torch.Tensor)
                   else np.asarray(pose_timestamps_ns)).astype(np.int64)
        if not np.any(ts_host == int(timestamp_ns)):                 # fail before any device work, like the reference
            raise ValueError(f"no pose at the sweep's timestamp {timestamp_ns}")
        table = PoseTable(ts_host, pose_quat_xyzw, pose_translation, dev)
    elif table.ts.device != _to_dev(np.zeros(0), torch.float64, dev).device:
        raise ValueError("the PoseTable lives on another device")
    row = table.row_at(timestamp_ns)
    pts = _to_dev(xyz, torch.float64, dev).reshape(-1, 3)
    off = _to_dev(offset_ns, torch.int64, dev).reshape(-1)
    if pts.shape[0] != off.shape[0]:
        raise ValueError("xyz and offset_ns disagree in length")
    tq = np.ascontiguousarray(table.quat_host[row])
    tt = np.ascontiguousarray(table.trans_host[row])
    out = torch.empty_like(pts)
    valid = torch.empty((pts.shape[0],), dtype=torch.uint8, device=dev)
    dropped = torch.zeros(1, dtype=torch.int32, device=dev)
    N.check(N.lib().rv3d_unmotion_compensate_table(ptr(pts), ptr(off), pts.shape[0], int(timestamp_ns), ptr(table.ts),
                                                   ptr(table.trans), ptr(table.intervals), table.n_poses, table.first_ns,
                                                   table.last_ns, tq.ctypes.data, tt.ctypes.data, ptr(out), ptr(valid),
                                                   ptr(dropped), stream_ptr(dev)), "rv3d_unmotion_compensate_table")
    keep = valid.view(torch.bool)
    if int(dropped.item()) == 0:          # the usual case (a sweep inside its log's pose table): nothing to compact
        return _back(out, xyz), _back(keep, xyz)
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


_LASER_TABLES = {}


def _laser_tables(dev, height: int, remap: bool):
    key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), int(height) == 32, bool(remap))
    t = _LASER_TABLES.get(key)
    if t is None:
        rows = torch.as_tensor((ROW_MAPPING_32 if height == 32 else ROW_MAPPING_64).astype(np.int64), device=dev)
        mapping = torch.as_tensor(LASER_MAPPING.astype(np.int64), device=dev) if remap else None
        t = _LASER_TABLES[key] = (rows, mapping)
    return t


def correct_laser_numbers(laser_numbers, log_id: str, height: int, log_ids: Optional[Collection[str]] = None,
                          device=None, validate: bool = True):
    """utils.py:211-226 -> row-mapped laser numbers (int64).  ``log_ids`` is the reference's ``LOG_IDS`` table
    (``datasets/argoverse/constants.py:269``: the logs recorded with the other beam ordering); it is dataset
    metadata, not arithmetic, and stays with the caller.  Out-of-table laser numbers raise ``IndexError`` like numpy;
    that check is the call's only host read -- ``validate=False`` skips it (out-of-table rows come back as -1) for
    callers that keep the stream asynchronous."""
    dev = _dev(device)
    las = _to_dev(laser_numbers, torch.int64, dev).reshape(-1)
    remap = log_ids is not None and log_id in log_ids
    rows, mapping = _laser_tables(dev, height, remap)
    out = torch.empty_like(las)
    bad = torch.zeros(1, dtype=torch.int32, device=dev) if validate else None
    N.check(N.lib().rv3d_correct_laser_numbers(ptr(las), las.shape[0], ptr(mapping), ptr(rows), rows.shape[0], ptr(out),
                                               ptr(bad), stream_ptr(dev)), "rv3d_correct_laser_numbers")
    if validate and int(bad.item()):
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
