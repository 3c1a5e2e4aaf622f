"""The rasterizer and the decoder replace libm's fp64 atan2 / exp / sqrt by shorter routines (csrc/fastmath.cuh) and the
scatter kernel decides a point's azimuth bin in float32, falling back to fp64 when it cannot PROVE that both paths
round to the same integer.  These tests measure the routines through the library's test hooks instead of trusting the
error analysis in the comments:

  * fast_atan2 / fast_exp / fast_sqrt against numpy's fp64 libm, in fp64 ulps (bar: <= 3 / 2 / 2; the results are
    rounded to float32 or half afterwards, where 1 fp64 ulp changes fewer than 1e-8 of the values);
  * atan2f_lite against fp64 atan2 of the same float32 arguments (bar: 5e-7 rad; the column band assumes 1e-6);
  * the float32 column decision: over > 1e7 points, including points placed within 1e-9 .. 1e-4 bins of a bin boundary
    in both column formulas and for even / odd bin counts, EVERY decided point must equal the reference's arithmetic
    with libm atan2, and so must the fp64 fallback's answer for the undecided ones.
"""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _fm(op, a, b=None):
    from rv3d import _native as N
    from rv3d._util import ptr, stream_ptr
    a = torch.as_tensor(a, dtype=torch.float64, device=DEV).contiguous()
    bt = None if b is None else torch.as_tensor(b, dtype=torch.float64, device=DEV).contiguous()
    out = torch.empty_like(a)
    N.check(N.lib().rv3d_debug_fastmath(op, ptr(a), ptr(bt) if bt is not None else None, ptr(out), a.numel(),
                                        stream_ptr(torch.device(DEV))), "rv3d_debug_fastmath")
    return out.cpu().numpy()


def _ulps(got, ref):
    return np.abs(got - ref) / np.spacing(np.abs(ref))


def test_fast_atan2_ulps():
    rng = np.random.default_rng(0)
    n = 4_000_000
    x = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-3, 3, n)
    y = rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-3, 3, n)
    y[::7] = x[::7] * (1 + rng.uniform(-5e-4, 5e-4, n)[::7])      # near the diagonal (swap boundary)
    y[::11] *= 1e-9                                                # tiny angles
    x[::13] = np.float32(x[::13])
    got = _fm(0, y, x)
    ref = np.arctan2(y, x)
    u = _ulps(got, ref)
    print(f"fast_atan2: max {u.max():.2f} ulp, mean {u.mean():.3f}")
    assert u.max() <= 3.0
    # special values go to libm
    sy = np.array([0.0, -0.0, 0.0, -0.0, 1.0, -1.0, np.inf, np.nan, 1.0, 1e-320, 1e200, 3.0])
    sx = np.array([1.0, 1.0, -1.0, -1.0, 0.0, -0.0, np.inf, 1.0, np.nan, 1e-320, -1e200, -0.0])
    g = _fm(0, sy, sx)
    r = np.arctan2(sy, sx)
    assert np.array_equal(np.isnan(g), np.isnan(r))
    ok = ~np.isnan(r)
    assert np.all(_ulps(g[ok], r[ok]) <= 2.0) and np.array_equal(np.signbit(g[ok]), np.signbit(r[ok]))


def test_fast_exp_and_sqrt_ulps():
    rng = np.random.default_rng(1)
    n = 4_000_000
    x = np.concatenate([rng.uniform(-12, 12, n), rng.uniform(-690, 690, n // 4), np.float32(rng.normal(0, 1, n // 4)).astype(np.float64)])
    u = _ulps(_fm(1, x), np.exp(x))
    print(f"fast_exp: max {u.max():.2f} ulp, mean {u.mean():.3f}")
    assert u.max() <= 2.0
    sp = np.array([0.0, -0.0, 710.0, -746.0, -800.0, np.inf, -np.inf, np.nan, 699.9999, -699.9999])
    g, r = _fm(1, sp), np.exp(sp)
    assert np.array_equal(np.isnan(g), np.isnan(r))
    ok = ~np.isnan(r) & np.isfinite(r) & (r > 0)
    assert np.all(_ulps(g[ok], r[ok]) <= 2.0) and np.array_equal(g[~ok & ~np.isnan(r)], r[~ok & ~np.isnan(r)])
    s = 10.0 ** rng.uniform(-50, 50, n)
    u = _ulps(_fm(2, s), np.sqrt(s))
    print(f"fast_sqrt: max {u.max():.2f} ulp")
    assert u.max() <= 2.0
    sp = np.array([0.0, 1e-300, 1e300, np.inf, 4.0])
    assert np.array_equal(_fm(2, sp), np.sqrt(sp))


def test_atan2f_lite_error():
    rng = np.random.default_rng(2)
    n = 8_000_000
    x = np.float32(rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-2, 2.5, n)).astype(np.float64)
    y = np.float32(rng.uniform(-1, 1, n) * 10.0 ** rng.uniform(-2, 2.5, n)).astype(np.float64)
    err = np.abs(_fm(3, y, x) - np.arctan2(y, x))
    print(f"atan2f_lite: max |err| {err.max():.3e} rad")
    assert err.max() < 5e-7


def _columns(points, az_bins, mode, offset):
    from rv3d import _native as N
    from rv3d._util import ptr, stream_ptr
    p = N.RasterParams()
    p.batch, p.max_points, p.height, p.width = 1, points.shape[0], 64, az_bins
    p.azimuth_bins, p.num_lasers, p.col_mode, p.reserved = az_bins, 64, mode, 0
    for k in range(3):
        p.lidar_offset[k] = float(offset[k])
    p.min_distance = 1.0
    pts = torch.as_tensor(points, dtype=torch.float32, device=DEV).contiguous()
    cf = torch.empty(points.shape[0], dtype=torch.int32, device=DEV)
    ce = torch.empty_like(cf)
    N.check(N.lib().rv3d_debug_column(p, ptr(pts), points.shape[0], ptr(cf), ptr(ce), stream_ptr(torch.device(DEV))),
            "rv3d_debug_column")
    return cf.cpu().numpy(), ce.cpu().numpy()


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("az_bins", [1800, 2650, 901, 3600])
def test_column_decision_is_never_wrong(mode, az_bins):
    rng = np.random.default_rng(100 * mode + az_bins)
    offset = np.array([1.356, 0.0, 1.726]) if az_bins != 901 else np.array([-1.43, 0.25, 2.184])
    n = 1_500_000
    # (a) ordinary returns; (b) points on a bin boundary +- 1e-9 .. 1e-4 bins; (c) points next to the sensor origin
    az = rng.uniform(-math.pi, math.pi, n)
    edge = (rng.integers(0, az_bins, n) + 0.5) * (math.tau / az_bins) - math.pi        # rounding boundaries of both formulas
    edge = edge + rng.choice([-1, 1], n) * 10.0 ** rng.uniform(-9, -4, n) * (math.tau / az_bins)
    half = rng.integers(0, az_bins, n) * (math.tau / az_bins) - math.pi + rng.normal(0, 1e-7, n)
    az = np.where(rng.uniform(0, 1, n) < 0.4, edge, az)
    az = np.where(rng.uniform(0, 1, n) < 0.1, half, az)
    r = 10.0 ** rng.uniform(-0.3, 2.2, n)
    r[::50] = 10.0 ** rng.uniform(-9, -2, n)[::50]
    pts = np.zeros((n, 4), np.float32)
    pts[:, 0] = r * np.cos(az) + offset[0]
    pts[:, 1] = r * np.sin(az) + offset[1]
    pts[::1000, :2] = np.float32(offset[:2])                                           # exactly on the axis of the sensor
    cf, ce = _columns(pts, az_bins, mode, offset)
    decided = cf >= 0
    wrong = decided & (cf != ce)
    fb_wrong = ~decided & ((-1 - cf) != ce)
    print(f"mode {mode} bins {az_bins}: {decided.mean() * 100:.2f} % decided in float32, "
          f"{int(wrong.sum())} wrong, fallback wrong {int(fb_wrong.sum())}")
    assert not wrong.any(), pts[wrong][:5]
    assert not fb_wrong.any(), pts[fb_wrong][:5]
    assert decided[:: 3].mean() > 0.3          # the fast path is actually in use
