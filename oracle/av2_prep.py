"""CPU restatement of the exporter's sweep preparation.  TEST INFRASTRUCTURE ONLY.

Paths relative to /root/reference.  Plain numpy float64, the polars plumbing replaced by array slicing.
scipy.spatial.transform.{Rotation, Slerp} (a dependency of the reference, scipy >= 1.11) is restated from its
published algorithm; tests/golden/prep.npz pins this file against the reference's own functions executed
verbatim with the real scipy (tests/golden/make_golden_prep.py).  av2.geometry.se3.SE3 is NOT installed here:
inverse() / transform_point_cloud() follow its published definition -> that one piece is "parity unpinned".
"""
from __future__ import annotations

import math

import numpy as np

from .rv_oracle import build_range_view_coordinates_converter, cart_to_sph, z_buffer

# datasets/argoverse/constants.py:231-266 and :453-488 (dataset tables)
LASER_MAPPING = np.array([4, 15, 0, 14, 6, 11, 2, 8, 10, 7, 12, 9, 5, 3, 13, 26, 1, 19, 30, 24, 18, 23, 28, 20, 22,
                          25, 16, 27, 21, 29, 17, 31])
ROW_MAPPING_32 = np.array([29, 15, 25, 18, 31, 19, 27, 22, 24, 20, 23, 26, 21, 17, 28, 30, 5, 1, 11, 14, 8, 3, 7, 10,
                           12, 6, 16, 4, 9, 2, 13, 0])


# ----- scipy Rotation, scalar-last quaternions ------------------------------------------------
def _normalize(q):
    return q / np.linalg.norm(q, axis=-1, keepdims=True)


def _compose(p, q):
    """Rotation p * q: apply q, then p."""
    pv, qv = p[..., :3], q[..., :3]
    v = p[..., 3:4] * qv + q[..., 3:4] * pv + np.cross(pv, qv)
    w = p[..., 3] * q[..., 3] - np.sum(pv * qv, axis=-1)
    return np.concatenate([v, w[..., None]], axis=-1)


def _inv(q):
    return np.concatenate([-q[..., :3], q[..., 3:4]], axis=-1)


def _as_rotvec(q):
    q = np.where(q[..., 3:4] < 0, -q, q)
    nv = np.linalg.norm(q[..., :3], axis=-1)
    angle = 2.0 * np.arctan2(nv, q[..., 3])
    small = angle <= 1e-3
    a2 = angle * angle
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(small, 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0, angle / np.sin(angle / 2.0))
    return scale[..., None] * q[..., :3]


def _from_rotvec(rv):
    n = np.linalg.norm(rv, axis=-1)
    small = n <= 1e-3
    n2 = n * n
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = np.where(small, 0.5 - n2 / 48.0 + n2 * n2 / 3840.0, np.sin(n / 2.0) / n)
    return np.concatenate([scale[..., None] * rv, np.cos(n / 2.0)[..., None]], axis=-1)


def quat_to_matrix(q):
    q = _normalize(np.asarray(q, dtype=np.float64))
    x, y, z, w = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    m = np.empty(q.shape[:-1] + (3, 3))
    m[..., 0, 0] = x * x - y * y - z * z + w * w
    m[..., 0, 1] = 2 * (x * y - z * w)
    m[..., 0, 2] = 2 * (x * z + y * w)
    m[..., 1, 0] = 2 * (x * y + z * w)
    m[..., 1, 1] = -x * x + y * y - z * z + w * w
    m[..., 1, 2] = 2 * (y * z - x * w)
    m[..., 2, 0] = 2 * (x * z - y * w)
    m[..., 2, 1] = 2 * (y * z + x * w)
    m[..., 2, 2] = -x * x - y * y + z * z + w * w
    return m


def slerp_matrices(pose_ts, pose_quat, times):
    """scipy Slerp(times=pose_ts, Rotation.from_quat(pose_quat))(times).as_matrix(): float64 timestamps, search side
    'left' minus one, t_min mapped to interval 0, rotvec of the interval scaled by alpha."""
    tt = np.asarray(pose_ts, dtype=np.float64)
    ct = np.asarray(times, dtype=np.float64)
    q = _normalize(np.asarray(pose_quat, dtype=np.float64))
    rotvecs = _as_rotvec(_normalize(_compose(_inv(q[:-1]), q[1:])))
    ind = np.searchsorted(tt, ct) - 1
    ind[ct == tt[0]] = 0
    if np.any((ind < 0) | (ind > len(q) - 2)):
        raise ValueError("Interpolation times must be within the pose range")
    alpha = (ct - tt[ind]) / np.diff(tt)[ind]
    return quat_to_matrix(_normalize(_compose(q[ind], _from_rotvec(rotvecs[ind] * alpha[:, None]))))


def unmotion_compensate(xyz, offset_ns, timestamp_ns, pose_ts, pose_quat, pose_t):
    """converters/av2/utils.py:229-295 -> (xyz_p (N',3), keep (N,) bool)."""
    pose_ts = np.asarray(pose_ts, dtype=np.int64)
    pose_t = np.asarray(pose_t, dtype=np.float64)
    t = int(timestamp_ns) + np.asarray(offset_ns).astype(np.int64)                     # :236
    keep = (t > pose_ts.min()) & (t < pose_ts.max())                                   # :237-240
    t, xyz = t[keep], np.asarray(xyz, dtype=np.float64)[keep]
    idx = np.searchsorted(pose_ts, t, side="left")                                     # :244
    ts_low, ts_high = pose_ts[idx - 1], pose_ts[idx]                                   # :247-251
    t_low, t_high = pose_t[idx - 1], pose_t[idx]                                       # :253-254
    rot = slerp_matrices(pose_ts, pose_quat, t)                                        # :255
    hit = np.nonzero(pose_ts == int(timestamp_ns))[0]                                  # :257-266
    if hit.size == 0:
        raise ValueError("no pose at the sweep timestamp")
    city_se3_roll = np.eye(4)
    city_se3_roll[:3, :3] = quat_to_matrix(np.asarray(pose_quat, dtype=np.float64)[hit[0]])
    city_se3_roll[:3, 3] = pose_t[hit[0]]
    city_se3_laser = np.eye(4)[None].repeat(len(rot), axis=0)
    city_se3_laser[:, :3, :3] = rot
    alpha = ((t - ts_low) / (ts_high - ts_low))[:, None]                               # :275
    city_se3_laser[:, :3, 3] = t_low * alpha + (1 - alpha) * t_high                    # :276-277 (weights as written)
    rot_inv = city_se3_laser[:, :3, :3].transpose(0, 2, 1)                             # :280-282
    t_inv = np.einsum("bij,bj->bi", rot_inv, -city_se3_laser[:, :3, 3])
    laser_se3_city = np.zeros_like(city_se3_laser)
    laser_se3_city[:, :3, :3] = rot_inv
    laser_se3_city[:, :3, 3] = t_inv
    # NB the reference leaves laser_se3_city[:, 3, 3] at 0 (np.zeros_like, :284): the homogeneous row of the
    # product is [0 0 0 0]... and is never read (:293 keeps the first three coordinates)
    laser_se3_roll = np.einsum("bij,jk->bik", laser_se3_city, city_se3_roll)           # :287
    xyz_hom = np.ones((len(xyz), 4))
    xyz_hom[:, :3] = xyz
    return np.einsum("bi,bji->bj", xyz_hom, laser_se3_roll)[:, :3], keep               # :289-293


def sensor_from_egovehicle(xyz, rotation, translation):
    """utils.py:54-57 with av2 SE3: inverse() = (R^T, R^T.(-t)); transform_point_cloud(p) = p @ R'^T + t'."""
    rotation = np.asarray(rotation, dtype=np.float64).reshape(3, 3)
    r_inv = rotation.T
    t_inv = r_inv.dot(-np.asarray(translation, dtype=np.float64))
    return np.asarray(xyz, dtype=np.float64) @ r_inv.T + t_inv


def correct_laser_numbers(laser_numbers, remap: bool, height: int):
    """utils.py:211-226 (`remap` = log_id in LOG_IDS).  Mutates a copy."""
    l = np.array(laser_numbers, dtype=np.int64, copy=True)
    if remap:
        l[l >= 32] = LASER_MAPPING[l[l >= 32] - 32] + 32
        l[l < 32] = LASER_MAPPING[l[l < 32]]
    return (ROW_MAPPING_32 if height == 32 else ROW_MAPPING_64_)[l]


from .rv_oracle import ROW_MAPPING_64 as ROW_MAPPING_64_  # noqa: E402


def build_range_view(cart, features, laser_number, offset_ns, rotation, translation, height, width,
                     build_uniform_inclination=False, return_winner=False):
    """utils.py:32-105 -> (8,H,W) f32 [x,y,z,intensity,laser_number,is_within_roi,timedelta_ns,range]."""
    cart_lidar = sensor_from_egovehicle(cart, rotation, translation)                    # :43-57
    laser_number = np.asarray(laser_number).astype(int)                                 # :59-61
    timestamp_ns = np.asarray(offset_ns).astype(float)                                  # :63
    sph = cart_to_sph(cart_lidar)                                                       # :68
    hybrid = build_range_view_coordinates_converter(cart_lidar, sph, laser_number, np.arange(height), height, width,
                                                    build_uniform_inclination)          # :69-77
    indices = hybrid[:, :2].astype(int).T                                               # :79
    distances = hybrid[:, -1]                                                           # :80
    feats = np.concatenate((np.asarray(features, dtype=np.float64), timestamp_ns[:, None], distances[:, None]), axis=-1).T
    return z_buffer(indices, distances, feats, height, width, return_winner=return_winner)   # :90
