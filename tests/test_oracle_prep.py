"""The oracle's restatement of the exporter's sweep preparation (oracle/av2_prep.py) against vectors minted from the
reference's own functions run verbatim with the real scipy (tests/golden/make_golden_prep.py)."""
import numpy as np
import pytest

from oracle import av2_prep
from tests.conftest import GOLDEN


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "prep.npz")


@pytest.mark.parametrize("tag", ["mid", "end"])
def test_unmotion_compensate_matches_reference(g, tag):
    xyz_p, keep = av2_prep.unmotion_compensate(g[f"um_{tag}_xyz"], g[f"um_{tag}_offset_ns"], int(g[f"um_{tag}_timestamp_ns"]),
                                               g["pose_ts"], g["pose_quat"], g["pose_trans"])
    assert np.array_equal(np.nonzero(keep)[0], g[f"um_{tag}_kept_rows"])          # the same rows survive the filter
    if tag == "end":
        assert 0 < keep.sum() < keep.size
    # float64 quaternion algebra in a different operation order than scipy's: agreement to ~1e-12 m at |p| <= 200 m
    np.testing.assert_allclose(xyz_p, g[f"um_{tag}_xyz_p"], rtol=0, atol=1e-10)


def test_slerp_matches_scipy_directly():
    from scipy.spatial.transform import Rotation, Slerp
    from tests import synth
    ts, quat, _ = synth.make_pose_table(50, seed=3)
    rng = np.random.default_rng(0)
    t = np.sort(rng.integers(ts[0], ts[-1], size=2000))
    t[:3] = [ts[0], ts[7], ts[-1]]                      # knots, including both ends
    ref = Slerp(ts, Rotation.from_quat(quat))(t).as_matrix()
    np.testing.assert_allclose(av2_prep.slerp_matrices(ts, quat, t), ref, rtol=0, atol=1e-13)


def test_correct_laser_numbers_matches_reference(g):
    assert np.array_equal(av2_prep.correct_laser_numbers(g["laser64"], False, 64), g["rows64_plain"])
    assert np.array_equal(av2_prep.correct_laser_numbers(g["laser64"], True, 64), g["rows64_remap"])
    assert np.array_equal(av2_prep.correct_laser_numbers(g["laser32"], False, 32), g["rows32_plain"])
    assert np.array_equal(av2_prep.correct_laser_numbers(g["laser32"], True, 32), g["rows32_remap"])
    assert not np.array_equal(g["rows64_plain"], g["rows64_remap"])


@pytest.mark.parametrize("uniform", [False, True])
def test_converter_build_range_view_matches_reference(g, uniform):
    img = av2_prep.build_range_view(g["brv_cart"], g["brv_features"], g["brv_laser"], g["brv_offset_ns"], g["brv_rotation"],
                                    g["brv_ext_trans"], 32, 1800, build_uniform_inclination=uniform)
    ref = g[f"brv_image_{int(uniform)}"]
    assert img.dtype == np.float32 and img.shape == ref.shape
    assert np.array_equal(img.view(np.uint32), ref.view(np.uint32))
    assert (ref[7] > 0).sum() > 3000
