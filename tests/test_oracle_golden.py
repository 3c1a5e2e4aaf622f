"""The oracle restatement vs golden vectors minted from the reference's own functions
(tests/golden/make_golden.py).  CPU only."""
import math

import numpy as np
import pytest
import torch

import oracle
from tests.conftest import GOLDEN


def _load(name):
    return np.load(GOLDEN / name)


@pytest.mark.parametrize("name", ["raster_a.npz", "raster_b.npz", "raster_adv.npz", "raster_f32.npz"])
def test_rasterize_matches_reference(name):
    g = _load(name)
    H, W = int(g["H"]), int(g["W"])
    img = oracle.build_range_view(g["xyz"], g["intensity"], g["laser"], g["mapping"], g["offset"],
                                  num_lasers=H, width=W, n_azimuth_bins=W)
    # same machine, same numpy ufuncs -> bit-exact, every channel
    assert img.dtype == np.float32 and img.shape == g["image"].shape
    assert np.array_equal(img.view(np.uint32), g["image"].view(np.uint32))


def test_rasterize_converter_flavour():
    g = _load("raster_conv.npz")
    H, W = int(g["H"]), int(g["W"])
    cart = g["xyz"].astype(np.float64)
    sph = oracle.cart_to_sph(cart)
    rng = sph[:, 2].copy()
    hyb = oracle.build_range_view_coordinates_converter(cart, sph, g["laser"].astype(int), np.arange(H), H, W)
    feats = np.concatenate([cart, g["intensity"][:, None].astype(np.float64),
                            g["laser"][:, None].astype(np.float64), rng[:, None]], axis=-1).T
    img = oracle.z_buffer(hyb[:, :2].astype(int).T, hyb[:, -1], feats, H, W)
    assert np.array_equal(img.view(np.uint32), g["image"].view(np.uint32))


@pytest.mark.parametrize("name", ["decode_a.npz", "decode_b.npz"])
def test_decode_matches_reference(name):
    g = _load(name)
    reg, cart = torch.from_numpy(g["regressands"]), torch.from_numpy(g["cart"])
    for flag in (True, False):
        out = oracle.decode_range_view(reg, cart, flag).numpy()
        np.testing.assert_allclose(out, g[f"cuboids_{int(flag)}"], rtol=1e-6, atol=1e-6)
    scores = torch.from_numpy(g["logits"]).sigmoid() * torch.from_numpy(g["mask"])
    scores, cats = scores.max(dim=1, keepdim=True)
    s, c, b = oracle.sample_by_range(scores, cats, torch.from_numpy(g["cuboids_1"]), cart,
                                     (0, 15, 30), (15, 30, math.inf), (8, 2, 1))
    assert np.array_equal(s.numpy(), g["sbr_scores"])
    assert np.array_equal(c.numpy(), g["sbr_categories"])
    assert np.array_equal(b.numpy(), g["sbr_cuboids"])


@pytest.mark.parametrize("mode", ["hard", "weighted"])
def test_nms_control_flow_matches_reference(mode):
    g = _load(f"nms_{mode}.npz")
    o = oracle.batched_multiclass_nms(torch.from_numpy(g["cuboids"]), torch.from_numpy(g["scores"]),
                                      torch.from_numpy(g["categories"]), 500, 20, 0.3, 0.1, mode)
    assert np.array_equal(o[0].numpy(), g["out_cuboids"])
    assert np.array_equal(o[1].numpy(), g["out_scores"])
    assert np.array_equal(o[2].numpy(), g["out_categories"])
    assert np.array_equal(o[3].numpy(), g["out_batch_index"])
    assert o[2].dtype == torch.float32 and o[3].dtype == torch.float32     # nms.py:51,242


@pytest.mark.parametrize("mode", ["hard", "weighted", "nonms"])
def test_pipeline_matches_reference(mode):
    g = _load(f"pipeline_{mode}.npz")
    head = {k: torch.from_numpy(g[k]) for k in ("logits", "regressands", "cart", "mask")}
    ms = {1: {"cart": head["cart"], "mask": head["mask"], 0: {"logits": head["logits"], "regressands": head["regressands"]}}}
    pp = {"num_pre_nms": 50000, "num_post_nms": 1000, "nms_threshold": 0.3, "min_confidence": 0.1,
          "nms_mode": "HARD" if mode == "nonms" else mode.upper()}
    p, s, c, b = oracle.range_decoder_decode(ms, pp, {0: ["A", "B", "C"]}, True, mode != "nonms",
                                             [0, 15, 30], [15, 30, math.inf], [8, 2, 1], use_nms=mode != "nonms")
    assert p.shape == g["params"].shape
    np.testing.assert_allclose(p.numpy(), g["params"], rtol=1e-6, atol=1e-6)
    assert np.array_equal(s.numpy(), g["scores"])
    assert np.array_equal(c.numpy(), g["categories"])
    assert np.array_equal(b.numpy(), g["batch_index"])


def test_empty_and_bad_mode():
    cub = torch.zeros(2, 10, 7)
    sc = torch.zeros(2, 10)
    ca = torch.zeros(2, 10, dtype=torch.int64)
    o = oracle.batched_multiclass_nms(cub, sc, ca, 10, 10, 0.3, 0.1, "hard")
    assert [tuple(x.shape) for x in o] == [(0, 7), (0, 1), (0, 1), (0, 1)]      # nms.py:250-253
    with pytest.raises(NotImplementedError):
        oracle.batched_multiclass_nms(cub, sc, ca, 10, 10, 0.3, 0.1, "soft")


def test_subsample_range_view_matches_reference():
    g = _load("subsample.npz")
    feats, cart, mask = (torch.from_numpy(g[k]) for k in ("features", "cart", "mask"))
    for ds, stride, mode in (("av2", 1, "circular"), ("av2", 4, "circular"), ("waymo", 4, "constant")):
        f, m, c = oracle.subsample_range_view(feats.clone(), mask.clone(), cart.clone(), ds, stride, mode)
        assert np.array_equal(f.numpy(), g[f"{ds}_{stride}_{mode}_f"])
        assert np.array_equal(m.numpy(), g[f"{ds}_{stride}_{mode}_m"])
        assert np.array_equal(c.numpy(), g[f"{ds}_{stride}_{mode}_c"])


def test_libm_hypot_restatement():
    """The restatement of glibc's hypot the CUDA rasterizer uses near float32 decision points == numpy.hypot, bit for
    bit, on the value ranges the path meets (differences of float32 coordinates and float64 sensor offsets), including
    the exits for y == 0 and |y| << |x|."""
    from oracle.libm_hypot import libm_hypot
    rng = np.random.default_rng(5)
    n = 2_000_000
    x = rng.uniform(-250, 250, n).astype(np.float32).astype(np.float64) - 1.356
    y = rng.uniform(-250, 250, n).astype(np.float32).astype(np.float64)
    z = rng.uniform(-30, 30, n).astype(np.float32).astype(np.float64) - 1.726
    y[:1000] = 0.0; x[1000:2000] = 0.0; y[2000:3000] = x[2000:3000]; y[3000:4000] *= 1e-17; x[4000:5000] *= 1e-3
    h = np.hypot(x, y)
    assert np.array_equal(libm_hypot(x, y).view(np.uint64), h.view(np.uint64))
    assert np.array_equal(libm_hypot(h, z).view(np.uint64), np.hypot(h, z).view(np.uint64))
    # and it is NOT simply sqrt(x^2 + y^2): the correction step matters
    assert (np.sqrt(x * x + y * y) != h).mean() > 0.05
