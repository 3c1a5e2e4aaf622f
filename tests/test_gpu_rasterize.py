"""CUDA rasterizer (through the C ABI) vs the golden vectors minted from the verbatim reference
and vs the CPU oracle on seeded inputs.  Bar: winner map, x/y/z/intensity AND range channels bit-exact (the radius
is libm's hypot restated wherever a last bit could matter); az / inc within 1 float32 ulp (device atan2 vs host libm
in fp64, then one cast)."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _ulp_diff(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    ia = a.view(np.int32).astype(np.int64)
    ib = b.view(np.int32).astype(np.int64)
    ia = np.where(ia < 0, -(ia & 0x7FFFFFFF), ia)
    ib = np.where(ib < 0, -(ib & 0x7FFFFFFF), ib)
    return np.abs(ia - ib)


def _run(xyz, inten, laser, mapping, offset, H, W, nbins=None, col_mode="library", min_distance=1.0):
    from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
    dev = torch.device("cuda:0")
    pts, las, cnt = pack_sweeps([(xyz, inten, laser)], dev)
    img, win = rasterize_sweeps(pts.to(dev), las.to(dev), cnt.to(dev),
                                torch.as_tensor(np.asarray(mapping, dtype=np.int32), device=dev), offset,
                                height=H, width=W, n_azimuth_bins=nbins, col_mode=col_mode,
                                min_distance=min_distance, return_winner=True)
    return img[0].cpu().numpy(), win[0].cpu().numpy()


def _check_image(img, ref):
    assert img.shape == ref.shape and img.dtype == np.float32
    # x, y, z, intensity: copies of the winning point -> the pixel assignment is bit-exact
    assert np.array_equal(img[3:].view(np.uint32), ref[3:].view(np.uint32))
    assert np.array_equal(img[2].view(np.uint32), ref[2].view(np.uint32)), \
        f"range channel: {(img[2] != ref[2]).sum()} pixels differ"
    d = _ulp_diff(img[:2], ref[:2])
    assert d.max() <= 1, f"az/inc differ by {d.max()} ulp"
    return int((d > 0).sum())


@pytest.mark.parametrize("name", ["raster_a.npz", "raster_b.npz", "raster_adv.npz"])
def test_golden(name):
    g = np.load(GOLDEN / name)
    H, W = int(g["H"]), int(g["W"])
    img, _ = _run(g["xyz"], g["intensity"], g["laser"], g["mapping"], g["offset"], H, W)
    _check_image(img, g["image"])


@pytest.mark.parametrize("shape", [("av2", 100_000, 64, 1800), ("waymo", 180_000, 64, 2650)])
@pytest.mark.parametrize("seed", [0, 1])
def test_full_size_vs_oracle(shape, seed):
    name, n, H, W = shape
    mapping = oracle.ROW_MAPPING_64 if name == "av2" else np.arange(H)
    xyz, inten, laser = synth.make_points(n, H, seed, extra_laser_frac=0.01)
    ref_img, ref_win = oracle.build_range_view(xyz, inten, laser, mapping, synth.LIDAR_OFFSET, num_lasers=H,
                                               width=W, n_azimuth_bins=W, return_winner=True)
    img, win = _run(xyz, inten, laser, mapping, synth.LIDAR_OFFSET, H, W)
    assert np.array_equal(win, ref_win)          # bit-exact pixel assignment
    _check_image(img, ref_img)


def test_zbuffer_f32_distances_golden():
    """The all-float32 flavour (float32 offset -> float32 distances and features, numpy/conversions.py:118-127 with a
    float32 `distances`): rv3d_zbuffer's float32 path against the golden image minted from the verbatim reference."""
    from rv3d.math.numpy.conversions import z_buffer
    g = np.load(GOLDEN / "raster_f32.npz")
    H, W = int(g["H"]), int(g["W"])
    xyz, off = g["xyz"], g["offset"]
    assert off.dtype == np.float32
    cart = xyz - off                                                       # float32 arithmetic, like the reference
    sph = oracle.cart_to_sph(cart)
    assert sph.dtype == np.float32
    feats = np.concatenate([sph, xyz, g["intensity"].reshape(-1, 1)], axis=1).T.copy()
    hyb = oracle.build_range_view_coordinates(cart, sph, g["laser"].astype(np.int64), g["mapping"], H, W)
    idx = np.ascontiguousarray(hyb[:, :2].T.astype(int))
    dist = np.ascontiguousarray(hyb[:, 2])
    assert dist.dtype == np.float32 and feats.dtype == np.float32
    img, win = z_buffer(idx, dist, feats, H, W, return_winner=True)
    ref_img, ref_win = oracle.z_buffer(idx, dist, feats, H, W, return_winner=True)
    assert np.array_equal(np.asarray(win), ref_win)
    assert np.array_equal(np.asarray(img).view(np.uint32), g["image"].view(np.uint32))


def test_radius_decision_points_are_exact():
    """Radii engineered to sit within a few fp64 ulps of a float32 value / of a midpoint between two float32 values /
    of min_distance: there sqrt(x^2 + y^2 + z^2) and the reference's hypot(hypot(x, y), z) round differently, and the
    winner (class bit of the z-buffer key), the range channel and the min_distance cut depend on the last bit."""
    rng = np.random.default_rng(11)
    off = np.array([1.356, 0.0, 1.726])
    n = 400_000
    # many points per pixel with nearly equal radii: one laser row, a few columns
    az = rng.uniform(0.30, 0.31, n)
    r = np.float32(rng.uniform(1.0, 60.0, n))
    r[: n // 4] = np.float32(1.0) + np.float32(rng.integers(-3, 4, n // 4)) * np.spacing(np.float32(1.0))
    xyz = np.stack([r * np.cos(az), r * np.sin(az), 0.05 * r], 1) + off
    xyz = xyz.astype(np.float32)
    inten = rng.random(n).astype(np.float32)
    laser = np.zeros(n, np.uint8)
    ref_img, ref_win = oracle.build_range_view(xyz, inten, laser, np.arange(4), off, num_lasers=4, width=1800,
                                               n_azimuth_bins=1800, return_winner=True)
    img, win = _run(xyz, inten, laser, np.arange(4), off, 4, 1800)
    assert np.array_equal(win, ref_win)
    _check_image(img, ref_img)
    # the sweep really contains decision points where the two radius formulas differ after the float32 cast
    c = xyz.astype(np.float64) - off
    fast = np.sqrt(np.sqrt(c[:, 0] ** 2 + c[:, 1] ** 2) ** 2 + c[:, 2] ** 2)
    ref = np.hypot(np.hypot(c[:, 0], c[:, 1]), c[:, 2])
    assert (fast != ref).mean() > 0.05


def test_h2_order_dependence():
    """Several f64 radii rounding to one float32: the serial loop's winner is order dependent
    (SURVEY H2); the packed key must reproduce it for every permutation."""
    xyz, inten, laser, off = synth.make_h2_points()
    H = int(laser.max()) + 1
    rng = np.random.default_rng(0)
    for _ in range(4):
        perm = rng.permutation(len(xyz))
        x, i, l = xyz[perm], inten[perm], laser[perm]
        ref_img, ref_win = oracle.build_range_view(x, i, l, np.arange(H), off, num_lasers=H, width=64,
                                                   n_azimuth_bins=64, return_winner=True)
        img, win = _run(x, i, l, np.arange(H), off, H, 64)
        assert np.array_equal(win, ref_win)
        _check_image(img, ref_img)


def test_converter_column_formula_and_reference_width_quirk():
    xyz, inten, laser = synth.make_points(50_000, 32, 9, offset=np.zeros(3))
    cart = xyz.astype(np.float64)
    sph = oracle.cart_to_sph(cart)
    hyb = oracle.build_range_view_coordinates_converter(cart, sph, laser.astype(int), np.arange(32), 32, 1200)
    _, ref_win = oracle.z_buffer(hyb[:, :2].astype(int).T, hyb[:, 2], cart.T, 32, 1200, return_winner=True)
    _, win = _run(xyz, inten, laser, np.arange(32), np.zeros(3), 32, 1200, col_mode="converter")
    assert np.array_equal(win, ref_win)
    # library wrapper quirk: columns for 1800 bins, ravelled with width 2650 (range_view.py:34-43)
    ref_img, ref_win = oracle.build_range_view(xyz, inten, laser, np.arange(32), np.zeros(3), num_lasers=32,
                                               width=2650, n_azimuth_bins=1800, return_winner=True)
    img, win = _run(xyz, inten, laser, np.arange(32), np.zeros(3), 32, 2650, nbins=1800)
    assert np.array_equal(win, ref_win)
    _check_image(img, ref_img)


def test_empty_and_ragged_batch():
    from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
    dev = torch.device("cuda:0")
    sweeps = [synth.make_points(n, 16, s) for s, n in enumerate([5000, 1, 777])]
    sweeps.insert(1, (np.zeros((0, 3), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.uint8)))
    pts, las, cnt = pack_sweeps(sweeps, dev)
    mapping = torch.arange(16, dtype=torch.int32, device=dev)
    img, win = rasterize_sweeps(pts.to(dev), las.to(dev), cnt.to(dev), mapping, synth.LIDAR_OFFSET, height=16,
                                width=300, return_winner=True)
    for b, (xyz, inten, laser) in enumerate(sweeps):
        ref_img, ref_win = oracle.build_range_view(xyz, inten, laser, np.arange(16), synth.LIDAR_OFFSET,
                                                   num_lasers=16, width=300, n_azimuth_bins=300, return_winner=True)
        assert np.array_equal(win[b].cpu().numpy(), ref_win)
        _check_image(img[b].cpu().numpy(), ref_img)
    assert (win[1] == -1).all() and (img[1] == 0).all()


def test_build_range_view_signature():
    from rv3d.math.range_view import build_range_view
    xyz, inten, laser = synth.make_points(30_000, 64, 4, extra_laser_frac=0.02)
    sweep = {"x": xyz[:, 0], "y": xyz[:, 1], "z": xyz[:, 2], "intensity": inten, "laser_number": laser}
    out = build_range_view(sweep, oracle.ROW_MAPPING_64, synth.LIDAR_OFFSET, num_lasers=64, width=1800)
    ref = oracle.build_range_view(xyz, inten, laser, oracle.ROW_MAPPING_64, synth.LIDAR_OFFSET, 64, 1800, 1800)
    assert isinstance(out, np.ndarray) and out.shape == (7, 64, 1800)
    _check_image(out, ref)


def test_idempotent_and_deterministic():
    xyz, inten, laser = synth.make_points(100_000, 64, 2)
    a = _run(xyz, inten, laser, np.arange(64), synth.LIDAR_OFFSET, 64, 1800)
    b = _run(xyz, inten, laser, np.arange(64), synth.LIDAR_OFFSET, 64, 1800)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
    # permutation invariance of the IMAGE (the winner index follows the permutation)
    perm = np.random.default_rng(0).permutation(len(xyz))
    c = _run(xyz[perm], inten[perm], laser[perm], np.arange(64), synth.LIDAR_OFFSET, 64, 1800)
    ref = oracle.build_range_view(xyz[perm], inten[perm], laser[perm], np.arange(64), synth.LIDAR_OFFSET, 64, 1800, 1800)
    _check_image(c[0], ref)


@pytest.mark.parametrize("W", [900, 1800, 3600])
def test_config4_multi_resolution(W):
    """BASELINE config 4: the same sweeps rasterized at 64 x {900, 1800, 3600}."""
    from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
    dev = torch.device("cuda:0")
    sweeps = [synth.make_points(100_000, 64, 40 + s) for s in range(4)]
    pts, las, cnt = pack_sweeps(sweeps, dev)
    mapping = torch.as_tensor(oracle.ROW_MAPPING_64.astype(np.int32), device=dev)
    img, win = rasterize_sweeps(pts.to(dev), las.to(dev), cnt.to(dev), mapping, synth.LIDAR_OFFSET, height=64,
                                width=W, return_winner=True)
    for b, (xyz, inten, laser) in enumerate(sweeps):
        ref_img, ref_win = oracle.build_range_view(xyz, inten, laser, oracle.ROW_MAPPING_64, synth.LIDAR_OFFSET,
                                                   num_lasers=64, width=W, n_azimuth_bins=W, return_winner=True)
        assert np.array_equal(win[b].cpu().numpy(), ref_win)
        _check_image(img[b].cpu().numpy(), ref_img)


def test_build_range_view_float64_columns():
    """The reference takes whatever dtype the frame holds (math/range_view.py:27-29): float64 x / y / z go through the
    free-standing device operators and match the oracle run on the same float64 values."""
    from rv3d.math.range_view import build_range_view
    H = 16
    xyz, inten, laser = synth.make_points(20_000, H, 21, extra_laser_frac=0.02)
    xyz64 = xyz.astype(np.float64) + 1e-9                                            # not representable in float32
    sweep = {"x": xyz64[:, 0], "y": xyz64[:, 1], "z": xyz64[:, 2], "intensity": inten, "laser_number": laser}
    got = build_range_view(sweep, np.arange(H), synth.LIDAR_OFFSET, num_lasers=H, width=1800)
    keep = laser < H
    ref = oracle.build_range_view(xyz64[keep], inten[keep], laser[keep], np.arange(H), synth.LIDAR_OFFSET, num_lasers=H,
                                  width=1800, n_azimuth_bins=1800)
    assert got.shape == ref.shape == (7, H, 1800) and got.dtype == np.float32
    assert np.array_equal(got[2:], ref[2:])                                          # range, x, y, z, intensity: bit-exact
    np.testing.assert_allclose(got[:2], ref[:2], rtol=0, atol=2.4e-7)               # az / inc: device vs host libm, 1 float32 ulp
