"""Golden vectors for the exporter's sweep preparation: run ONCE in the build container.

    python tests/golden/make_golden_prep.py        ->  tests/golden/prep.npz

The reference's own functions are executed VERBATIM from /root/reference/converters/av2/utils.py:
``unmotion_compensate`` (with the real scipy Rotation / Slerp), ``correct_laser_numbers`` (with the real
LOG_IDS / LASER_MAPPING / ROW_MAPPING tables) and ``build_range_view``.  Two absent dependencies are shimmed:

  polars  a ~80-line duck-typed stand-in below implementing exactly the frame calls those functions make
          (select / filter / with_columns / row gather / search_sorted / to_numpy); it carries no arithmetic.
  av2     ``SE3`` (inverse, transform_point_cloud) restated from its published definition -> the sensor-frame
          transform inside build_range_view is pinned only up to that restatement.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.golden import make_golden as mg  # noqa: E402  (installs the import stubs, adds the reference to sys.path)
from tests import synth  # noqa: E402

OUT = Path(__file__).resolve().parent


# ----------------------------------------------------------------------------------------------
# minimal polars stand-in
# ----------------------------------------------------------------------------------------------
class Series:
    def __init__(self, v):
        self.v = np.asarray(v)

    def min(self): return self.v.min()
    def max(self): return self.v.max()
    def cast(self, dt): return Series(self.v.astype(dt))
    def to_numpy(self): return self.v
    def __radd__(self, o): return Series(o + self.v)
    def __sub__(self, o): return Series(self.v - o)
    def __len__(self): return len(self.v)

    def search_sorted(self, other, side="any"):
        return Series(np.searchsorted(self.v, other.v, side="left" if side in ("any", "left") else "right").astype(np.int64))


class Pred:
    def __init__(self, fn): self.fn = fn
    def __and__(self, o): return Pred(lambda f: self.fn(f) & o.fn(f))


class Col:
    def __init__(self, names): self.names = names
    def __gt__(self, o): return Pred(lambda f: f.c[self.names[0]] > o)
    def __lt__(self, o): return Pred(lambda f: f.c[self.names[0]] < o)
    def __eq__(self, o): return Pred(lambda f: f.c[self.names[0]] == o)
    def eq(self, o): return self == o


class Lit:
    def __init__(self, v): self.v = np.asarray(v)


class Frame:
    def __init__(self, cols, schema=None):
        self.c = {k: np.asarray(v.v if isinstance(v, (Series, Lit)) else v) for k, v in cols.items()}

    def __getitem__(self, k):
        if isinstance(k, str):
            return Series(self.c[k])
        rows = k.v if isinstance(k, Series) else np.asarray(k)
        return Frame({n: a[rows] for n, a in self.c.items()})

    def with_columns(self, *args, **kw):
        new = dict(self.c)
        for k, v in kw.items():
            new[k] = v.v if isinstance(v, (Series, Lit)) else v
        return Frame(new)

    def filter(self, pred):
        m = pred.fn(self)
        return Frame({n: a[m] for n, a in self.c.items()})

    def select(self, *what):
        names = []
        for w in what:
            names += list(w.names) if isinstance(w, Col) else ([w] if isinstance(w, str) else list(w))
        return Frame({n: self.c[n] for n in names})

    def to_numpy(self):
        return np.stack([self.c[n] for n in self.c], axis=1)


class MiniPolars:
    Int64, Float32, UInt8, Boolean = np.int64, np.float32, np.uint8, np.bool_
    DataFrame = Frame

    @staticmethod
    def col(*names):
        flat = []
        for n in names:
            flat += [n] if isinstance(n, str) else list(n)
        return Col(flat)

    @staticmethod
    def lit(v): return Lit(v)


class MiniSE3:
    """av2.geometry.se3.SE3 (published definition): p' = p @ R^T + t; inverse = (R^T, R^T.(-t))."""

    def __init__(self, rotation, translation):
        self.rotation, self.translation = np.asarray(rotation, dtype=np.float64), np.asarray(translation, dtype=np.float64)

    def inverse(self):
        return MiniSE3(self.rotation.T, self.rotation.T.dot(-self.translation))

    def transform_point_cloud(self, pts):
        return pts @ self.rotation.T + self.translation


def main():
    from scipy.spatial.transform import Rotation, Slerp

    cu = mg.load_converter_utils()
    cu.pl = MiniPolars
    cu.SE3 = MiniSE3
    out = {}

    # ---------------- unmotion_compensate ----------------------------------------------------
    ts, quat, trans = synth.make_pose_table(300, seed=31)
    poses = Frame({"timestamp_ns": ts, "qx": quat[:, 0], "qy": quat[:, 1], "qz": quat[:, 2], "qw": quat[:, 3],
                   "tx_m": trans[:, 0], "ty_m": trans[:, 1], "tz_m": trans[:, 2]})
    slerp = Slerp(ts, Rotation.from_quat(quat))                      # converters/av2/export.py:61-64
    for tag, pose_row in (("mid", 140), ("end", 292)):               # "end": the sweep runs past the last pose -> rows dropped
        xyz, off, inten, laser, roi = synth.make_raw_sweep(6000, seed=40 + pose_row)
        off[:8] = [0, 1, 255, 256, 257, 99_999_999, 50_000_000, 50_000_001]
        sweep = Frame({"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2], "offset_ns": off.astype(np.int32),
                       "row": np.arange(len(off))})
        res = cu.unmotion_compensate(sweep, poses, int(ts[pose_row]), slerp)
        out.update({f"um_{tag}_xyz": xyz, f"um_{tag}_offset_ns": off, f"um_{tag}_timestamp_ns": np.int64(ts[pose_row]),
                    f"um_{tag}_kept_rows": res.c["row"],
                    f"um_{tag}_xyz_p": np.stack([res.c["x_p"], res.c["y_p"], res.c["z_p"]], 1)})
    out.update(pose_ts=ts, pose_quat=quat, pose_trans=trans)

    # ---------------- correct_laser_numbers ---------------------------------------------------
    rng = np.random.default_rng(5)
    l64 = rng.integers(0, 64, size=4000).astype(np.uint8)
    l32 = rng.integers(0, 32, size=4000).astype(np.uint8)
    remapped_log = cu.LOG_IDS[0]
    out.update(laser64=l64, laser32=l32,
               rows64_plain=cu.correct_laser_numbers(l64.copy().astype(np.int64), "not-a-listed-log", 64),
               rows64_remap=cu.correct_laser_numbers(l64.copy().astype(np.int64), remapped_log, 64),
               rows32_plain=cu.correct_laser_numbers(l32.copy().astype(np.int64), "not-a-listed-log", 32),
               rows32_remap=cu.correct_laser_numbers(l32.copy().astype(np.int64), remapped_log, 32))

    # ---------------- build_range_view (converter) --------------------------------------------
    xyz, off, inten, laser, roi = synth.make_raw_sweep(20000, seed=77)
    keep = laser < 32
    xyz, off, inten, laser, roi = xyz[keep], off[keep], inten[keep], laser[keep], roi[keep]
    ext_q = np.array([0.0012, -0.0031, 0.0052, 0.99998])              # egovehicle_SE3_up_lidar-like: ~1 deg off identity
    ext_t = np.array([1.35, 0.0, 1.64])
    extrinsics = Frame({"sensor_name": np.array(["up_lidar", "down_lidar"]), "tx_m": np.array([ext_t[0], 1.355]),
                        "ty_m": np.array([ext_t[1], 0.0]), "tz_m": np.array([ext_t[2], 1.565]),
                        "qx": np.array([ext_q[0], 0.0]), "qy": np.array([ext_q[1], 1.0]), "qz": np.array([ext_q[2], 0.0]),
                        "qw": np.array([ext_q[3], 0.0])})
    feats = np.stack([xyz[:, 0], xyz[:, 1], xyz[:, 2], np.round(inten * 255), laser.astype(np.float64), roi], 1)
    lidar = Frame({"x_p": xyz[:, 0], "y_p": xyz[:, 1], "z_p": xyz[:, 2], "laser_number": laser.astype(np.uint8),
                   "offset_ns": off.astype(np.int32)})
    for uniform in (False, True):
        frame = cu.build_range_view(lidar, extrinsics, feats, "up_lidar", 32, 1800, build_uniform_inclination=uniform)
        cols = ["x", "y", "z", "intensity", "laser_number", "is_within_roi", "timedelta_ns", "range"]
        out[f"brv_image_{int(uniform)}"] = np.stack([np.asarray(frame.c[c]) for c in cols], 0).reshape(8, 32, 1800)
    out.update(brv_cart=xyz, brv_features=feats, brv_laser=laser, brv_offset_ns=off, brv_ext_quat=ext_q, brv_ext_trans=ext_t,
               brv_rotation=Rotation.from_quat(ext_q).as_matrix())
    np.savez_compressed(OUT / "prep.npz", **out)
    print("written", OUT / "prep.npz", {k: getattr(v, "shape", None) for k, v in out.items()})


if __name__ == "__main__":
    main()
