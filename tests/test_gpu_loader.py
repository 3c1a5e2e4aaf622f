"""SURVEY 8f row 1 (the step between the rasterizer and the backbone): fused feature / cart / mask assembly +
subsample_range_view vs the restatement of prototype/loader.py:623-650, 792-815 (torch on CPU)."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _image(seed, H=32, W=600):
    xyz, inten, laser = synth.make_points(25_000, H, seed)
    return torch.from_numpy(oracle.build_range_view(xyz, inten, laser, np.arange(H), synth.LIDAR_OFFSET,
                                                    num_lasers=H, width=W, n_azimuth_bins=W))


def _reference_inputs(img, names, dataset, stride, mode):
    ch = {"azimuth": 0, "inclination": 1, "range": 2, "x": 3, "y": 4, "z": 5, "intensity": 6}
    feats = torch.stack([img[ch[n]].tanh() if (dataset == "waymo" and n == "intensity") else img[ch[n]] for n in names])
    cart = img[3:6].clone()
    mask = (img[2:3] > 0.0)
    return oracle.subsample_range_view(feats, mask, cart, dataset, stride, mode)


@pytest.mark.parametrize("dataset,stride,mode", [("av2", 1, "circular"), ("av2", 4, "circular"), ("waymo", 1, "circular"),
                                                 ("waymo", 4, "constant"), ("av2", 2, "constant")])
def test_fused_inputs(dataset, stride, mode):
    from rv3d.prototype.loader import range_view_inputs
    names = ("intensity", "range", "x", "y", "z")
    imgs = torch.stack([_image(s) for s in (0, 1)])
    f, m, c = range_view_inputs(imgs.to(DEV), names, dataset, stride, mode)
    for b in range(2):
        rf, rm, rc = _reference_inputs(imgs[b], names, dataset, stride, mode)
        assert f[b].shape == rf.shape and m[b].shape == rm.shape and c[b].shape == rc.shape
        assert torch.equal(m[b].cpu(), rm) and torch.equal(c[b].cpu(), rc)
        if dataset == "waymo":
            np.testing.assert_allclose(f[b].cpu().numpy(), rf.numpy(), rtol=1e-6, atol=1e-7)   # tanhf ulp
        else:
            assert torch.equal(f[b].cpu(), rf)


@pytest.mark.parametrize("dataset,stride,mode", [("av2", 1, "circular"), ("waymo", 4, "circular"), ("av2", 4, "constant")])
def test_subsample_range_view_drop_in(dataset, stride, mode):
    from rv3d.prototype.loader import subsample_range_view
    img = _image(3)
    feats, cart, mask = img[[6, 2, 3, 4, 5, 0]].clone(), img[3:6].clone(), img[2:3] > 0
    ref = oracle.subsample_range_view(feats.clone(), mask, cart, dataset, stride, mode)
    got = subsample_range_view(feats.to(DEV), mask.to(DEV), cart.to(DEV), dataset, stride, mode)
    for g, r in zip(got, ref):
        assert g.dtype == r.dtype and torch.equal(g.cpu(), r)


@pytest.mark.parametrize("dataset,stride,mode,names", [
    ("av2", 1, "circular", ("intensity", "range", "x", "y", "z")),
    ("waymo", 4, "circular", ("intensity", "range", "x", "y", "z")),
    ("waymo", 1, "constant", ("azimuth", "inclination", "range", "intensity")),
    ("av2", 2, "constant", ("range",)),
])
def test_rasterize_inputs_equals_rasterize_then_assemble(dataset, stride, mode, names):
    """rv3d_rasterize_inputs (raw sweeps -> network inputs, the 7-plane image never written) is bit-identical to
    rasterize_sweeps followed by range_view_inputs, whose two halves are pinned against the reference separately."""
    from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
    from rv3d.prototype.loader import range_view_inputs, rasterize_inputs
    H, W = 32, 600
    sweeps = [synth.make_points(n, H, s, extra_laser_frac=0.01) for s, n in ((5, 25_000), (6, 9_000), (7, 0))]
    sweeps[2] = (np.zeros((0, 3), np.float32), np.zeros((0,), np.float32), np.zeros((0,), np.uint8))
    pts, las, cnt = [t.to(DEV) for t in pack_sweeps(sweeps, DEV)]
    mapping = torch.arange(H, dtype=torch.int32, device=DEV)
    img = rasterize_sweeps(pts, las, cnt, mapping, synth.LIDAR_OFFSET, height=H, width=W)
    want = range_view_inputs(img, names, dataset, stride, mode)
    got = rasterize_inputs(pts, las, cnt, mapping, synth.LIDAR_OFFSET, height=H, width=W, feature_column_names=names,
                           dataset_name=dataset, x_stride=stride, mode=mode)
    for g, w_ in zip(got, want):
        assert g.shape == w_.shape and g.dtype == w_.dtype and torch.equal(g, w_)
    assert got[1][2].sum() == 0 and got[1][0].sum() > 0        # the empty sweep has an empty mask
