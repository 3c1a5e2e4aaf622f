"""Seeded synthetic workloads for the rasterize -> decode -> NMS path.

Used by tests/, tests/golden/make_golden.py and bench.py.  Everything is generated on the
host from ``numpy.random.default_rng(seed)`` / a seeded ``torch.Generator`` so the CPU
oracle and the CUDA path see identical bytes.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

LIDAR_OFFSET = np.array([1.356, 0.0, 1.726])  # datasets/argoverse/av2.py:162 (reference)

# (l, w, h) priors in metres, cycled over the class index
_PRIORS = np.array([[4.6, 2.0, 1.7], [0.8, 0.8, 1.75], [1.8, 0.7, 1.4], [10.5, 2.9, 3.3], [6.5, 2.5, 2.8]])


# --------------------------------------------------------------------------- #
# raw sweeps                                                                   #
# --------------------------------------------------------------------------- #
def make_points(n: int, n_lasers: int, seed: int, offset=LIDAR_OFFSET, extra_laser_frac: float = 0.0
                ) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """-> xyz (n,3) f32 ego frame, intensity (n,) f32, laser (n,) u8."""
    rng = np.random.default_rng(seed)
    laser = rng.integers(0, n_lasers, size=n)
    az = rng.uniform(-math.pi, math.pi, size=n)
    r = np.clip(rng.gamma(2.0, 12.0, size=n), 0.2, 200.0)
    near = rng.random(n) < 0.02
    r[near] = rng.uniform(0.2, 1.0, size=int(near.sum()))            # exercises the d < 1.0 skip
    inc = np.deg2rad(-25.0 + 40.0 * laser / max(n_lasers - 1, 1) + rng.normal(0.0, 0.05, size=n))
    xyz = np.stack([r * np.cos(inc) * np.cos(az), r * np.cos(inc) * np.sin(az), r * np.sin(inc)], axis=1)
    xyz = (xyz + np.asarray(offset, dtype=np.float64)).astype(np.float32)
    intensity = rng.random(n).astype(np.float32)
    if extra_laser_frac > 0:                                          # lasers >= n_lasers get filtered
        bad = rng.random(n) < extra_laser_frac
        laser[bad] = n_lasers + rng.integers(0, 8, size=int(bad.sum()))
    return xyz, intensity, laser.astype(np.uint8)


def make_adversarial_points(seed: int = 7, H: int = 8) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """Points engineered for the z-buffer / column edge cases (use with offset = 0, W = 64):
    exact duplicates (tie -> lowest index), radii identical after rounding, az = +-pi
    (y = +0.0 / -0.0), radii below / exactly at min_distance, many points per pixel."""
    rng = np.random.default_rng(seed)
    pts, las = [], []
    for row in range(H):
        base = np.float32(5.0 + row)
        # a run of adjacent float32 x values on the +x axis, shuffled, duplicated
        xs = base + np.arange(12, dtype=np.float32) * np.spacing(base)
        xs = np.concatenate([xs, xs[:4]])
        rng.shuffle(xs)
        for x in xs:
            pts.append((x, 0.0, 0.0)); las.append(row)
        # -x axis: az = +pi (y=+0) and -pi (y=-0)
        pts.append((-base, 0.0, 0.0)); las.append(row)
        pts.append((-base - 1, -0.0, 0.0)); las.append(row)
        # at / below min distance
        pts.append((1.0, 0.0, 0.0)); las.append(row)
        pts.append((0.5, 0.0, 0.0)); las.append(row)
        pts.append((np.nextafter(np.float32(1.0), np.float32(0.0)), 0.0, 0.0)); las.append(row)
        # random cloud confined to a few columns
        a = rng.uniform(0.3, 0.45, size=40)
        rr = rng.uniform(0.8, 30.0, size=40)
        for ai, ri in zip(a, rr):
            pts.append((ri * math.cos(ai), ri * math.sin(ai), 0.1 * ri)); las.append(row)
    xyz = np.asarray(pts, dtype=np.float32)
    perm = rng.permutation(len(xyz))
    xyz, las = xyz[perm], np.asarray(las, dtype=np.uint8)[perm]
    return xyz, rng.random(len(xyz)).astype(np.float32), las


def make_h2_points(seed: int = 3, groups: int = 64) -> Tuple[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """f32 coordinates + an f64 sensor offset chosen so that several f64 radii in one pixel
    round to the SAME float32 (the serial z-buffer winner is then order-dependent).
    Returns xyz, intensity, laser, offset; use H = groups, W = 64."""
    rng = np.random.default_rng(seed)
    offset = np.array([-4.1, 0.0, 0.0])          # cart.x = x + 4.1 : x in [4,8) (ulp 4.8e-7) -> r in [8,16) (ulp 9.5e-7)
    pts, las = [], []
    for g in range(groups):
        base = np.float32(rng.uniform(4.5, 7.0))
        xs = base + rng.integers(0, 10, size=14).astype(np.float32) * np.spacing(base)
        for x in xs:
            pts.append((x, 0.0, 0.0)); las.append(g % 256)
    xyz = np.asarray(pts, dtype=np.float32)
    return xyz, rng.random(len(xyz)).astype(np.float32), np.asarray(las, dtype=np.uint8), offset


# --------------------------------------------------------------------------- #
# dense head outputs                                                           #
# --------------------------------------------------------------------------- #
def _range_image_cart(B: int, H: int, W: int, rng: np.random.Generator, fill: float = 0.68):
    """A plausible (B,3,H,W) cart image in the ego frame + (B,1,H,W) validity mask: one return
    per pixel along the pixel's own ray, Gamma-distributed range, ``1-fill`` empty pixels."""
    az = math.pi - (np.arange(W) + 0.5) * (math.tau / W)                # col = W - az' convention
    inc = np.deg2rad(-25.0 + 40.0 * (H - 1 - np.arange(H)) / max(H - 1, 1))
    r = np.clip(rng.gamma(2.0, 12.0, size=(B, H, W)), 1.0, 200.0)
    valid = rng.random((B, H, W)) < fill
    ci, si = np.cos(inc)[None, :, None], np.sin(inc)[None, :, None]
    x = r * ci * np.cos(az)[None, None, :] + LIDAR_OFFSET[0]
    y = r * ci * np.sin(az)[None, None, :] + LIDAR_OFFSET[1]
    z = r * si + LIDAR_OFFSET[2]
    cart = np.stack([x, y, z], axis=1) * valid[:, None]
    return cart.astype(np.float32), valid[:, None]


def make_head_outputs(B: int, C: int, H: int, W: int, seed: int, n_objects: int = 96,
                      fp_rate: float = 0.25, distinct_scores: bool = True, min_conf: float = 0.1
                      ) -> Dict[str, torch.Tensor]:
    """-> {"logits" (B,C,H,W) f32, "regressands" (B,8,H,W) f32, "cart" (B,3,H,W) f32,
    "mask" (B,1,H,W) bool}.  Objects attract clusters of confident, mutually-overlapping
    boxes (the NMS work); a ``fp_rate`` fraction of background pixels fires on a random class."""
    rng = np.random.default_rng(seed)
    cart, mask = _range_image_cart(B, H, W, rng)
    logits = rng.normal(-6.0, 1.0, size=(B, C, H, W)).astype(np.float32)
    reg = np.empty((B, 8, H, W), dtype=np.float32)
    # background regressands: small offsets, class-prior-ish sizes, random heading
    bg_cls = rng.integers(0, C, size=(B, H, W))
    pri = _PRIORS[bg_cls % len(_PRIORS)]                                 # (B,H,W,3)
    reg[:, 0:3] = rng.normal(0.0, 0.5, size=(B, 3, H, W))
    reg[:, 3:6] = (np.log(pri) + rng.normal(0.0, 0.3, size=pri.shape)).transpose(0, 3, 1, 2)
    reg[:, 6:8] = rng.normal(0.0, 1.0, size=(B, 2, H, W))
    fire = (rng.random((B, H, W)) < fp_rate) & mask[:, 0]
    bi, hi, wi = np.nonzero(fire)
    logits[bi, bg_cls[fire], hi, wi] = rng.normal(0.0, 1.5, size=len(bi)).astype(np.float32)

    phi = np.arctan2(cart[:, 1].astype(np.float64), cart[:, 0].astype(np.float64))
    for b in range(B):
        ctr = np.concatenate([rng.uniform(-75, 75, size=(n_objects, 2)), rng.uniform(-1.0, 1.0, size=(n_objects, 1))], 1)
        cls = rng.integers(0, C, size=n_objects)
        lwh = _PRIORS[cls % len(_PRIORS)] * np.exp(rng.normal(0.0, 0.1, size=(n_objects, 3)))
        yaw = rng.uniform(-math.pi, math.pi, size=n_objects)
        px, py = cart[b, 0].astype(np.float64), cart[b, 1].astype(np.float64)
        for o in range(n_objects):
            dx, dy = px - ctr[o, 0], py - ctr[o, 1]
            rad = 0.75 * math.hypot(lwh[o, 0], lwh[o, 1])
            near = (np.abs(dx) < rad) & (np.abs(dy) < rad) & mask[b, 0]
            if not near.any():
                continue
            hh, ww = np.nonzero(near)
            c, s = math.cos(yaw[o]), math.sin(yaw[o])
            lx = c * dx[near] + s * dy[near]
            ly = -s * dx[near] + c * dy[near]
            inside = (np.abs(lx) < 0.75 * lwh[o, 0]) & (np.abs(ly) < 0.75 * lwh[o, 1])   # 1.5x dilated
            hh, ww = hh[inside], ww[inside]
            if len(hh) == 0:
                continue
            n = len(hh)
            ox = ctr[o, 0] - px[hh, ww]; oy = ctr[o, 1] - py[hh, ww]
            oz = ctr[o, 2] - cart[b, 2, hh, ww]
            p = phi[b, hh, ww]
            enc = np.stack([np.cos(p) * ox + np.sin(p) * oy, -np.sin(p) * ox + np.cos(p) * oy, oz,
                            np.full(n, math.log(lwh[o, 0])), np.full(n, math.log(lwh[o, 1])),
                            np.full(n, math.log(lwh[o, 2])), np.sin(yaw[o] - p), np.cos(yaw[o] - p)], axis=0)
            reg[b, :, hh, ww] = (enc + rng.normal(0.0, 0.15, size=enc.shape)).T.astype(np.float32)
            logits[b, cls[o], hh, ww] = rng.normal(1.0, 1.5, size=n).astype(np.float32)

    out = {"logits": torch.from_numpy(logits), "regressands": torch.from_numpy(reg),
           "cart": torch.from_numpy(cart), "mask": torch.from_numpy(mask)}
    if distinct_scores:
        _make_scores_distinct(out, min_conf, rng)
    return out


def _make_scores_distinct(head: Dict[str, torch.Tensor], min_conf: float, rng: np.random.Generator) -> None:
    """Nudge logits until no two surviving pixels of one (sweep, class) share a float32 score
    (top-k tie order is unspecified upstream, so goldens must not contain ties)."""
    logits = head["logits"]
    B = logits.shape[0]
    for _ in range(50):
        s = logits.sigmoid() * head["mask"]
        v, c = s.max(dim=1)
        dup_total = 0
        for b in range(B):
            vb, cb = v[b].flatten(), c[b].flatten()
            live = torch.nonzero(vb >= min_conf * 0.5).flatten()
            key = cb[live].double() * 4.0 + vb[live].double()
            order = torch.argsort(key, stable=True)
            ks = key[order]
            same = torch.nonzero(ks[1:] == ks[:-1]).flatten()
            if len(same) == 0:
                continue
            dup_total += len(same)
            pix = live[order[same + 1]]
            hh, ww = pix // logits.shape[3], pix % logits.shape[3]
            bump = torch.from_numpy(rng.uniform(1e-4, 1e-2, size=len(pix)).astype(np.float32))
            logits[b, cb[pix], hh, ww] -= bump
        if dup_total == 0:
            return
    raise RuntimeError("could not make scores distinct")


# --------------------------------------------------------------------------- #
# free-standing NMS candidates                                                 #
# --------------------------------------------------------------------------- #
def make_nms_candidates(B: int, K: int, n_classes: int, n_objects: int, seed: int, spread: float = 60.0,
                        frac_clustered: float = 0.8) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """-> cuboids (B,K,7) f32 [x,y,z,l,w,h,yaw], scores (B,K) f32 (distinct per sweep), categories (B,K) int64."""
    rng = np.random.default_rng(seed)
    cub = np.empty((B, K, 7), dtype=np.float32)
    cat = np.empty((B, K), dtype=np.int64)
    sc = np.empty((B, K), dtype=np.float32)
    for b in range(B):
        ctr = rng.uniform(-spread, spread, size=(n_objects, 2))
        ocl = rng.integers(0, n_classes, size=n_objects)
        oyaw = rng.uniform(-math.pi, math.pi, size=n_objects)
        olwh = _PRIORS[ocl % len(_PRIORS)] * np.exp(rng.normal(0, 0.1, size=(n_objects, 3)))
        which = rng.integers(0, n_objects, size=K)
        clustered = rng.random(K) < frac_clustered
        xy = np.where(clustered[:, None], ctr[which] + rng.normal(0, 0.35, size=(K, 2)),
                      rng.uniform(-spread, spread, size=(K, 2)))
        lwh = np.where(clustered[:, None], olwh[which] * np.exp(rng.normal(0, 0.08, size=(K, 3))),
                       _PRIORS[rng.integers(0, len(_PRIORS), size=K)] * np.exp(rng.normal(0, 0.3, size=(K, 3))))
        yaw = np.where(clustered, oyaw[which] + rng.normal(0, 0.1, size=K), rng.uniform(-math.pi, math.pi, size=K))
        cub[b, :, 0:2], cub[b, :, 2], cub[b, :, 3:6], cub[b, :, 6] = xy, rng.normal(0, 0.5, size=K), lwh, yaw
        cat[b] = np.where(clustered, ocl[which], rng.integers(0, n_classes, size=K))
        s = rng.permutation(K).astype(np.float64) / K                # distinct by construction
        sc[b] = (0.02 + 0.97 * s).astype(np.float32)
        assert len(np.unique(sc[b])) == K
    return torch.from_numpy(cub), torch.from_numpy(sc), torch.from_numpy(cat)


def make_pose_table(n_poses: int, seed: int, t0_ns: int = 315969904359876000, period_ns: int = 10_000_000):
    """A city_SE3_egovehicle-like log: -> (timestamps (M,) i64 strictly increasing, quat_xyzw (M,4) f64 (not exactly
    unit, like a feather file's rounded columns), translation (M,3) f64 in city coordinates).  A car driving a gentle
    curve at ~10 m/s with small roll / pitch."""
    rng = np.random.default_rng(seed)
    ts = t0_ns + np.arange(n_poses, dtype=np.int64) * period_ns + rng.integers(-400_000, 400_000, size=n_poses)
    ts = np.sort(ts)
    sec = (ts - ts[0]).astype(np.float64) * 1e-9
    yaw = 0.7 + 0.25 * sec + 0.05 * np.sin(1.3 * sec)
    pitch = 0.01 * np.sin(2.1 * sec) + rng.normal(0, 1e-4, n_poses)
    roll = 0.008 * np.cos(1.7 * sec) + rng.normal(0, 1e-4, n_poses)
    cy, sy, cp, sp, cr, sr = np.cos(yaw / 2), np.sin(yaw / 2), np.cos(pitch / 2), np.sin(pitch / 2), np.cos(roll / 2), np.sin(roll / 2)
    qw = cr * cp * cy + sr * sp * sy
    qx = sr * cp * cy - cr * sp * sy
    qy = cr * sp * cy + sr * cp * sy
    qz = cr * cp * sy - sr * sp * cy
    quat = np.stack([qx, qy, qz, qw], 1)
    quat = np.round(quat, 12)                                  # stored columns are not exactly unit length
    if n_poses > 8:
        quat[5] = -quat[5]                                     # the double cover shows up in real logs
    speed = 10.0
    x = 2000.0 + np.cumsum(np.r_[0.0, np.diff(sec)] * speed * np.cos(yaw))
    y = 1500.0 + np.cumsum(np.r_[0.0, np.diff(sec)] * speed * np.sin(yaw))
    z = 20.0 + 0.05 * np.sin(0.9 * sec)
    return ts, quat, np.stack([x, y, z], 1)


def make_raw_sweep(n: int, seed: int, sweep_ns: int = 100_000_000):
    """-> (xyz (N,3) f64 holding float32 values, like the exporter's cast of the feather columns,
    offset_ns (N,) i64 in [0, sweep_ns), intensity (N,) f64, laser_number (N,) i64 in [0, 64), roi (N,) f64)."""
    rng = np.random.default_rng(seed)
    xyz32, inten, laser = make_points(n, 64, seed, offset=LIDAR_OFFSET)
    off = rng.integers(0, sweep_ns, size=n).astype(np.int64)
    roi = (rng.random(n) < 0.8).astype(np.float64)
    return xyz32.astype(np.float64), off, inten.astype(np.float64), laser.astype(np.int64), roi


def make_assignment_inputs(B: int, C: int, H: int, W: int, seed: int, n_instances: int = 12):
    """Training-time inputs of compute_classification_targets: predictions `input` and encoded `target` (B,8,H,W) f32,
    labels (B,H,W) i64 in [0, C] (C = background), cart, mask (B,1,H,W) bool, panoptics (B,1,H,W) i64 (0 = background).
    Instances are rectangular pixel patches; predictions are the targets plus noise so the BEV IoUs spread over (0, 1)."""
    rng = np.random.default_rng(seed)
    cart, mask = _range_image_cart(B, H, W, rng)
    target = np.zeros((B, 8, H, W), dtype=np.float32)
    target[:, 0:3] = rng.normal(0.0, 0.8, size=(B, 3, H, W))
    pri = _PRIORS[rng.integers(0, len(_PRIORS), size=(B, H, W))]
    target[:, 3:6] = np.log(pri).transpose(0, 3, 1, 2)
    ang = rng.uniform(-math.pi, math.pi, size=(B, H, W))
    target[:, 6], target[:, 7] = np.sin(ang), np.cos(ang)
    inp = target + rng.normal(0.0, 0.25, size=target.shape).astype(np.float32)
    inp[:, 3:6] = target[:, 3:6] + rng.normal(0.0, 0.1, size=(B, 3, H, W)).astype(np.float32)
    labels = np.full((B, H, W), C, dtype=np.int64)
    pan = np.zeros((B, 1, H, W), dtype=np.int64)
    for b in range(B):
        for inst in range(1, n_instances + 1):
            if inst == 3:
                continue                                            # an id that never appears (one_hot column of zeros)
            h0, w0 = int(rng.integers(0, H - 2)), int(rng.integers(0, W - 12))
            hh, ww = int(rng.integers(1, 4)), int(rng.integers(1, 12))
            pan[b, 0, h0:h0 + hh, w0:w0 + ww] = inst
            labels[b, h0:h0 + hh, w0:w0 + ww] = int(rng.integers(0, C))
    pan[~mask] = 0
    t = torch.from_numpy
    return {"input": t(inp), "target": t(target), "labels": t(labels), "cart": t(cart), "mask": t(mask),
            "panoptics": t(pan)}
