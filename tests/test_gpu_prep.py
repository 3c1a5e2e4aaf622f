"""Sweep preparation in front of the rasterizer through the C ABI: CUDA vs the golden vectors
minted from the reference's own functions and vs the oracle at production sizes.  Bars: row filter, laser rows and
the pixel assignment exact; float64 coordinates within 1e-9 m (the device composes the same SE(3) maps in a
different, better conditioned order than the reference's 4x4 products -- differences are ~1e-12 m)."""
import numpy as np
import pytest
import torch

from oracle import av2_prep
from tests import synth
from tests.conftest import GOLDEN

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
ATOL = 1e-9


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "prep.npz")


@pytest.mark.parametrize("tag", ["mid", "end"])
def test_unmotion_compensate_golden(g, tag):
    from rv3d.converters.av2.utils import unmotion_compensate
    xyz_p, keep = unmotion_compensate(g[f"um_{tag}_xyz"], g[f"um_{tag}_offset_ns"], int(g[f"um_{tag}_timestamp_ns"]),
                                      g["pose_ts"], g["pose_quat"], g["pose_trans"], device=DEV)
    assert isinstance(xyz_p, np.ndarray) and xyz_p.dtype == np.float64 and keep.dtype == np.bool_
    assert np.array_equal(np.nonzero(keep)[0], g[f"um_{tag}_kept_rows"])
    np.testing.assert_allclose(xyz_p, g[f"um_{tag}_xyz_p"], rtol=0, atol=ATOL)
    assert np.abs(xyz_p - g[f"um_{tag}_xyz_p"]).max() < 1e-10


def test_unmotion_compensate_full_size_vs_oracle():
    from rv3d.converters.av2.utils import unmotion_compensate
    ts, quat, trans = synth.make_pose_table(3000, seed=9)
    xyz, off, *_ = synth.make_raw_sweep(200_000, seed=10)
    t0 = int(ts[1500])
    ref, keep_ref = av2_prep.unmotion_compensate(xyz, off, t0, ts, quat, trans)
    out, keep = unmotion_compensate(torch.from_numpy(xyz).to(DEV), torch.from_numpy(off).to(DEV), t0, ts, quat, trans)
    assert out.is_cuda and out.dtype == torch.float64
    assert np.array_equal(keep.cpu().numpy(), keep_ref)
    np.testing.assert_allclose(out.cpu().numpy(), ref, rtol=0, atol=ATOL)
    # the correction is real: centimetres to decimetres at 10 m/s over a 100 ms sweep
    assert np.abs(ref - xyz[keep_ref]).max() > 0.05


def test_unmotion_compensate_pose_table_and_irregular_poses():
    """One resident PoseTable serves many sweeps; a table with very uneven spacing defeats the interpolation guess and
    takes the binary-search fallback; both against the oracle (and the per-call form)."""
    from rv3d.converters.av2.utils import PoseTable, unmotion_compensate
    ts, quat, trans = synth.make_pose_table(600, seed=21)
    rng = np.random.default_rng(3)
    keep_rows = np.sort(rng.choice(np.arange(1, 599), size=120, replace=False))
    keep_rows = np.concatenate([[0], keep_rows[keep_rows < 80], keep_rows[keep_rows > 400], [599]])   # a 3 s hole + jitter
    for sel in (slice(None), keep_rows):
        t_s, q_s, p_s = ts[sel], quat[sel], trans[sel]
        table = PoseTable(t_s, q_s, p_s, device=DEV)
        for row in (5, len(t_s) // 2, len(t_s) - 4):
            t0 = int(t_s[row])
            xyz, off, *_ = synth.make_raw_sweep(20_000, seed=30 + row)
            off = off.copy()
            off[:50] = rng.integers(-(t0 - int(t_s[0])) - 5, int(t_s[-1]) - t0 + 5, size=50)    # anywhere in (and just outside) the table
            if row != 5:                                  # rows the reference's filter drops: the compaction branch
                off[50], off[51] = int(t_s[0]) - t0, int(t_s[-1]) - t0 + 3
            ref, keep_ref = av2_prep.unmotion_compensate(xyz, off, t0, t_s, q_s, p_s)
            out, keep = unmotion_compensate(xyz, off, t0, table)
            assert np.array_equal(keep, keep_ref) and (row == 5 or not keep[50:52].any())
            np.testing.assert_allclose(out, ref, rtol=0, atol=ATOL)
            out2, keep2 = unmotion_compensate(xyz, off, t0, t_s, q_s, p_s, device=DEV)
            assert np.array_equal(out, out2) and np.array_equal(keep, keep2)
    with pytest.raises(ValueError):
        unmotion_compensate(xyz, off, int(ts[7]) + 3, PoseTable(ts, quat, trans, device=DEV))


def test_unmotion_compensate_edges():
    from rv3d.converters.av2.utils import unmotion_compensate
    ts, quat, trans = synth.make_pose_table(16, seed=2)
    with pytest.raises(ValueError):                      # no pose carries the sweep's timestamp
        unmotion_compensate(np.zeros((4, 3)), np.zeros(4, dtype=np.int64), int(ts[3]) + 1, ts, quat, trans, device=DEV)
    out, keep = unmotion_compensate(np.zeros((0, 3)), np.zeros(0, dtype=np.int64), int(ts[3]), ts, quat, trans, device=DEV)
    assert out.shape == (0, 3) and keep.shape == (0,)
    # times exactly ON the first / last pose are dropped (strict inequalities), one ns inside is kept
    off = np.array([ts[0] - ts[3], ts[0] - ts[3] + 1, ts[-1] - ts[3] - 1, ts[-1] - ts[3]], dtype=np.int64)
    pts = np.arange(12, dtype=np.float64).reshape(4, 3)
    out, keep = unmotion_compensate(pts, off, int(ts[3]), ts, quat, trans, device=DEV)
    ref, keep_ref = av2_prep.unmotion_compensate(pts, off, int(ts[3]), ts, quat, trans)
    assert keep.tolist() == [False, True, True, False] == keep_ref.tolist()
    np.testing.assert_allclose(out, ref, rtol=0, atol=ATOL)
    # reference quirk kept (utils.py:276 puts alpha on the LOWER pose): a point stamped exactly at the sweep's own pose
    # gets the PREVIOUS pose's translation, so it does not come back unchanged -- in the reference either
    out, _ = unmotion_compensate(pts[:1], np.zeros(1, dtype=np.int64), int(ts[3]), ts, quat, trans, device=DEV)
    ref, _ = av2_prep.unmotion_compensate(pts[:1], np.zeros(1, dtype=np.int64), int(ts[3]), ts, quat, trans)
    np.testing.assert_allclose(out, ref, rtol=0, atol=ATOL)
    assert np.abs(ref - pts[:1]).max() > 0.05


def test_sensor_from_egovehicle_vs_oracle(g):
    from rv3d.converters.av2.utils import sensor_from_egovehicle
    xyz = synth.make_raw_sweep(100_000, seed=4)[0]
    out = sensor_from_egovehicle(xyz, g["brv_rotation"], g["brv_ext_trans"], device=DEV)
    np.testing.assert_allclose(out, av2_prep.sensor_from_egovehicle(xyz, g["brv_rotation"], g["brv_ext_trans"]), rtol=0, atol=1e-12)
    assert sensor_from_egovehicle(np.zeros((0, 3)), np.eye(3), np.zeros(3), device=DEV).shape == (0, 3)


def test_correct_laser_numbers_golden(g):
    from rv3d.converters.av2.utils import correct_laser_numbers
    listed = ("some-log", "another-log")
    for lasers, h in ((g["laser64"], 64), (g["laser32"], 32)):
        plain = correct_laser_numbers(lasers.astype(np.int64), "unlisted", h, log_ids=listed, device=DEV)
        remap = correct_laser_numbers(lasers.astype(np.int64), "some-log", h, log_ids=listed, device=DEV)
        assert plain.dtype == np.int64
        assert np.array_equal(plain, g[f"rows{h}_plain"]) and np.array_equal(remap, g[f"rows{h}_remap"])
    assert np.array_equal(correct_laser_numbers(g["laser64"].astype(np.int64), "x", 64, device=DEV), g["rows64_plain"])
    with pytest.raises(IndexError):                      # a 64-beam number against the 32-row table, like numpy
        correct_laser_numbers(np.array([3, 40]), "x", 32, device=DEV)
    lazy = correct_laser_numbers(np.array([3, 40, 7]), "x", 32, device=DEV, validate=False)    # no host read: -1 marks the row
    assert lazy[1] == -1 and np.array_equal(lazy[[0, 2]], correct_laser_numbers(np.array([3, 7]), "x", 32, device=DEV))
    odd = g["laser64"].astype(np.int64)[1:-2]            # odd length, 8-byte-aligned start: the scalar tail / unaligned form
    assert np.array_equal(correct_laser_numbers(torch.from_numpy(g["laser64"].astype(np.int64)).to(DEV)[1:-2], "x", 64).cpu().numpy(),
                          g["rows64_plain"][1:-2]) and len(odd) > 10


@pytest.mark.parametrize("uniform", [False, True])
def test_converter_build_range_view_golden(g, uniform):
    from rv3d.converters.av2.utils import build_range_view
    img, winner = build_range_view(g["brv_cart"], g["brv_features"], g["brv_laser"], g["brv_offset_ns"], g["brv_rotation"],
                                   g["brv_ext_trans"], 32, 1800, build_uniform_inclination=uniform, device=DEV,
                                   return_winner=True)
    ref = g[f"brv_image_{int(uniform)}"]
    _, ref_winner = av2_prep.build_range_view(g["brv_cart"], g["brv_features"], g["brv_laser"], g["brv_offset_ns"],
                                              g["brv_rotation"], g["brv_ext_trans"], 32, 1800, uniform, return_winner=True)
    assert img.dtype == np.float32 and img.shape == ref.shape
    assert np.array_equal(winner, ref_winner)                                    # pixel assignment: exact
    assert np.array_equal(img[:7].view(np.uint32), ref[:7].view(np.uint32))      # carried features: exact
    np.testing.assert_allclose(img[7], ref[7], rtol=1.2e-7, atol=0)              # range: <= 1 float32 ulp
    assert (img[7].view(np.uint32) == ref[7].view(np.uint32)).mean() > 0.9999
