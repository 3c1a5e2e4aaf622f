"""Training-time callers through the C ABI vs the golden vectors of the verbatim reference module
and vs the oracle.  Bars: foreground / background / regression masks exact; rotated IoU bit-exact (same routine as
NMS); affinities within 1e-6 (BEV: IoU of boxes decoded in fp64 on both sides; Gaussian: float32 exp)."""
import numpy as np
import pytest
import torch

from oracle import assign_oracle
from tests import synth
from tests.conftest import GOLDEN
from tests.test_oracle_assign import CFGS

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN / "assign.npz")


def _run(d, cfg, bg):
    from rv3d.math.ops.assignment import compute_classification_targets
    dv = {k: v.to(DEV) for k, v in d.items()}
    return compute_classification_targets(dv["input"], dv["target"], dv["labels"], dv["cart"], cfg, dv["mask"], dv["panoptics"], bg)


@pytest.mark.parametrize("tag", list(CFGS))
def test_compute_classification_targets_golden(g, tag):
    d = {k: torch.from_numpy(g[k]) for k in ("input", "target", "labels", "cart", "mask", "panoptics")}
    aff, fg, bgm, rw = _run(d, CFGS[tag], 3)
    assert aff.shape == g[f"{tag}_affinities"].shape and aff.dtype == torch.float32
    assert fg.dtype == torch.float32 and bgm.dtype == torch.bool and rw.dtype == torch.bool
    assert np.array_equal(fg.cpu().numpy(), g[f"{tag}_foreground"])
    assert np.array_equal(bgm.cpu().numpy(), g[f"{tag}_background"])
    assert np.array_equal(rw.cpu().numpy(), g[f"{tag}_reg_weights"])
    np.testing.assert_allclose(aff.cpu().numpy(), g[f"{tag}_affinities"], rtol=1e-5, atol=1e-6)


def test_compute_classification_targets_full_size_vs_oracle():
    d = synth.make_assignment_inputs(4, 3, 64, 1024, seed=5, n_instances=60)
    cfg = dict(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=8, normalize_affinities=False, sigma=1.0)
    ref = assign_oracle.compute_classification_targets(d["input"], d["target"], d["labels"], d["cart"], cfg, d["mask"], d["panoptics"], 3)
    got = _run(d, cfg, 3)
    for r, o in zip(ref[1:], got[1:]):
        assert torch.equal(r, o.cpu())
    np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-6)
    assert int(ref[1].sum()) > 200


@pytest.mark.parametrize("cfg", [
    dict(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=8, normalize_affinities=False, sigma=1.0),
    dict(affinity_fn="bev", enable_azimuth_invariant_targets=False, k=1, normalize_affinities=False, sigma=1.0),
    dict(affinity_fn="gaussian", enable_azimuth_invariant_targets=True, k=3, normalize_affinities=False, sigma=0.75),
    dict(affinity_fn="gaussian", enable_azimuth_invariant_targets=True, k=float("inf"), normalize_affinities=False, sigma=0.75),
    dict(affinity_fn="gaussian", enable_azimuth_invariant_targets=True, k=64, normalize_affinities=False, sigma=0.75),
])
def test_fused_targets_equal_composed(cfg):
    """The fused call (foreground-only decode, atomic top-k slot lists) against the same function composed from the
    free-standing operators (dense decodes + one stable sort): masks identical, affinities to float32 rounding of the
    Gaussian (the BEV IoU is the same routine: bit-equal).  Large instances (hundreds of pixels >> k) contend for the slots."""
    from rv3d.math.ops import assignment as A
    d = synth.make_assignment_inputs(3, 3, 64, 512, seed=11, n_instances=40)
    pan = d["panoptics"]
    pan[0, 0, 10:40, 100:160] = 41                        # one 1800-pixel instance
    pan[~d["mask"]] = 0
    dv = {k: v.to(DEV) for k, v in d.items()}
    args = (dv["input"], dv["target"], dv["labels"], dv["cart"], cfg, dv["mask"], dv["panoptics"], 3)
    fused = A.compute_classification_targets(*args)
    hinted = A.compute_classification_targets(*args, max_instances=64)
    comp = A._compute_classification_targets_composed(*args[:4], dict(cfg), *args[5:], str(cfg["affinity_fn"]).upper(), A._k_slots(cfg["k"]))
    for f, h in zip(fused, hinted):
        assert torch.equal(f, h)
    for f, c in zip(fused[1:], comp[1:]):
        assert f.dtype == c.dtype and torch.equal(f, c)
    if str(cfg["affinity_fn"]).upper() == "BEV":
        assert torch.equal(fused[0], comp[0])
    else:
        torch.testing.assert_close(fused[0], comp[0], rtol=2e-6, atol=1e-7)
    assert int(fused[1].sum()) > 100


def test_compute_classification_targets_edges():
    d = synth.make_assignment_inputs(1, 2, 8, 64, seed=6, n_instances=4)
    cfg = dict(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=4, normalize_affinities=False, sigma=1.0)
    none = dict(d, panoptics=torch.zeros_like(d["panoptics"]))                 # no instance at all
    aff, fg, bgm, _ = _run(none, cfg, 2)
    assert float(aff.abs().sum()) == 0 and float(fg.sum()) == 0 and torch.equal(bgm.cpu(), d["mask"])
    with pytest.raises(UnboundLocalError):                                     # the reference's own failure mode
        _run(d, dict(cfg, normalize_affinities=True), 2)
    with pytest.raises(NotImplementedError):
        _run(d, dict(cfg, affinity_fn="l2"), 2)
    k0 = _run(d, dict(cfg, k=0), 2)                                            # topk(0): nothing is foreground
    assert float(k0[1].sum()) == 0


def test_pair_affinities_golden(g):
    from rv3d.math.ops.assignment import iou_2d_axis_aligned, iou_3d_axis_aligned
    a, b = torch.from_numpy(g["pair_a"]).to(DEV), torch.from_numpy(g["pair_b"]).to(DEV)
    assert np.array_equal(iou_2d_axis_aligned(a, b, normalize_affinities=False).cpu().numpy(), g["pair_iou2d"])   # bit-exact IoU
    np.testing.assert_allclose(iou_3d_axis_aligned(a, b, normalize_affinities=False).cpu().numpy(), g["pair_iou3d"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(iou_3d_axis_aligned(a, b, normalize_affinities=True).cpu().numpy(), g["pair_iou3d_norm"], rtol=1e-6, atol=1e-7)


def test_box_iou_rotated_mmcv_published_vector():
    """mmcv's own unit-test vector for box_iou_rotated (upstream tests/test_ops/test_box_iou_rotated.py, atol 1e-4) through
    the device operator: pins the stand-in's rotation direction (see tests/test_oracle_iou.py::test_mmcv_published_vectors)."""
    from rv3d.math.ops.assignment import box_iou_rotated
    b1 = torch.tensor([[1.0, 1.0, 3.0, 4.0, 0.5], [2.0, 2.0, 3.0, 4.0, 0.6], [7.0, 7.0, 8.0, 8.0, 0.4]], device=DEV)
    b2 = torch.tensor([[0.0, 2.0, 2.0, 5.0, 0.3], [2.0, 1.0, 3.0, 3.0, 0.5], [5.0, 5.0, 6.0, 7.0, 0.4]], device=DEV)
    want = np.array([[0.3708, 0.4351, 0.0000], [0.1104, 0.4487, 0.0424], [0.0000, 0.0000, 0.3622]], np.float32)
    assert np.allclose(box_iou_rotated(b1, b2).cpu().numpy(), want, atol=1e-4)
    assert np.allclose(box_iou_rotated(b1, b2, aligned=True).cpu().numpy(), np.diag(want), atol=1e-4)


def test_box_iou_rotated_all_pairs_and_collision_test():
    from rv3d.math.ops.assignment import box_iou_rotated
    from rv3d.prototype.loader import intersection_test
    cub = synth.make_nms_candidates(1, 300, 1, 20, seed=3)[0][0]
    a, b = cub[:120], cub[120:]
    ref = assign_oracle.box_iou_rotated(a[:, [0, 1, 3, 4, 6]], b[:, [0, 1, 3, 4, 6]])
    got = box_iou_rotated(a[:, [0, 1, 3, 4, 6]].to(DEV), b[:, [0, 1, 3, 4, 6]].to(DEV))
    assert got.shape == (120, 180) and np.array_equal(got.cpu().numpy(), ref.numpy())
    assert (ref > 0).sum() > 50
    assert np.array_equal(intersection_test(a.to(DEV), b.to(DEV)).cpu().numpy(), ref.numpy())
    assert box_iou_rotated(a[:0, :5].to(DEV), b[:, :5].to(DEV)).shape == (0, 180)
    with pytest.raises(ValueError):
        box_iou_rotated(a[:3, :5].to(DEV), b[:4, :5].to(DEV), aligned=True)
