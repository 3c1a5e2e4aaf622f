"""CUDA decode kernels (through the C ABI) vs golden vectors from the verbatim reference and vs the
CPU oracle.  Bars: candidate index set and categories exact; box parameters within 1e-5 relative
(north_star); scores within 1e-6 relative (float32 sigmoid differs by an ulp between libms)."""
import math

import numpy as np
import pytest
import torch

import oracle
from tests import synth
from tests.conftest import GOLDEN
from tests.util import PP, SBR, ms_outputs, to_dev, unpack_candidates

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
RTOL = 1e-5      # north_star: decoded box parameters within 1e-5 relative


def _close(a, b, rtol=RTOL, atol=1e-6):
    np.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize("name", ["decode_a.npz", "decode_b.npz"])
def test_decode_range_view_golden(name):
    from rv3d.math.ops.coding import decode_range_view
    g = np.load(GOLDEN / name)
    reg, cart = torch.from_numpy(g["regressands"]).to(DEV), torch.from_numpy(g["cart"]).to(DEV)
    for flag in (True, False):
        out = decode_range_view(reg, cart, flag)
        assert out.dtype == torch.float32 and tuple(out.shape) == g[f"cuboids_{int(flag)}"].shape
        _close(out.cpu().numpy(), g[f"cuboids_{int(flag)}"])
        # in practice the fp64 path reproduces the reference bit for bit almost everywhere
        same = (out.cpu().numpy().view(np.uint32) == g[f"cuboids_{int(flag)}"].view(np.uint32)).mean()
        assert same > 0.999


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_decode_range_view_half(dtype):
    from rv3d.math.ops.coding import decode_range_view
    head = synth.make_head_outputs(1, 3, 8, 128, 5, n_objects=4)
    reg, cart = head["regressands"].to(dtype), head["cart"].to(dtype)
    ref = oracle.decode_range_view(reg, cart, True)
    out = decode_range_view(reg.to(DEV), cart.to(DEV), True)
    assert out.dtype == dtype
    assert torch.equal(out.cpu(), ref)      # both round an fp64 result once to the 16-bit type


@pytest.mark.parametrize("name", ["decode_a.npz", "decode_b.npz"])
def test_sample_by_range_golden(name):
    from rv3d.nn.decoders.range_decoder import sample_by_range
    g = np.load(GOLDEN / name)
    s, c, b = sample_by_range(torch.from_numpy(g["scores"]).to(DEV), torch.from_numpy(g["categories"]).to(DEV),
                              torch.from_numpy(g["cuboids_1"]).to(DEV), torch.from_numpy(g["cart"]).to(DEV),
                              (0, 15, 30), (15, 30, math.inf), (8, 2, 1))
    assert np.array_equal(s.cpu().numpy(), g["sbr_scores"])
    assert np.array_equal(c.cpu().numpy(), g["sbr_categories"])
    assert np.array_equal(b.cpu().numpy(), g["sbr_cuboids"])


def _check_candidates(head, sbr: bool, az_inv: bool = True, min_conf: float = 0.1):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    C = head["logits"].shape[1]
    dec = RangeDecoder(az_inv, sbr, *SBR)
    pp = dict(PP, min_confidence=min_conf)
    tasks = {0: [f"c{i}" for i in range(C)]}
    cand = dec.candidates(ms_outputs(to_dev(head, DEV)), pp, tasks)
    n = cand.count()
    got = unpack_candidates(cand, n)
    params, scores, cats = oracle.range_decoder_decode(ms_outputs(head), pp, tasks, az_inv, sbr, *SBR,
                                                       return_candidates=True)
    scores, cats, params = scores.numpy(), cats.numpy(), params.numpy()
    thr = np.float32(min_conf)
    live = scores >= thr
    borderline = np.abs(scores - thr) <= 2e-7
    ref_set = {(b, k) for b, k in zip(*np.nonzero(live))}
    got_set = set(zip(got["sweep"].tolist(), got["k"].tolist()))
    diff = ref_set ^ got_set
    assert all(borderline[b, k] for b, k in diff), f"{len(diff)} candidates differ away from the threshold"
    assert len(got_set) == n, "duplicate candidate rows"
    b_, k_ = got["sweep"], got["k"]
    ok = ~borderline[b_, k_]
    np.testing.assert_allclose(got["score"][ok], scores[b_, k_][ok], rtol=1e-6, atol=0)
    assert np.array_equal(got["category"][ok], cats[b_, k_][ok])
    _close(got["boxes"][ok], params[b_, k_][ok])
    return n


@pytest.mark.parametrize("name", ["decode_a.npz", "decode_b.npz"])
@pytest.mark.parametrize("sbr", [True, False])
def test_fused_candidates_golden_inputs(name, sbr):
    g = np.load(GOLDEN / name)
    head = {k: torch.from_numpy(g[k]) for k in ("logits", "regressands", "cart", "mask")}
    assert _check_candidates(head, sbr) > 0


@pytest.mark.parametrize("shape", [(2, 26, 64, 1800), (2, 3, 64, 2650)])
def test_fused_candidates_full_size(shape):
    B, C, H, W = shape
    head = synth.make_head_outputs(B, C, H, W, seed=3, n_objects=64, distinct_scores=False)
    n = _check_candidates(head, True)
    assert n > 0.05 * B * H * W


def test_fused_candidates_edge_cases():
    head = synth.make_head_outputs(2, 4, 7, 131, seed=2, n_objects=5)          # HW % 4 != 0 -> scalar path
    _check_candidates(head, True)
    _check_candidates(head, False, az_inv=False)
    # saturated logits: many classes round to sigmoid == 1.0 -> torch.max picks the FIRST
    head = synth.make_head_outputs(1, 5, 8, 64, seed=4, n_objects=3)
    head["logits"][:, 1:4] = torch.tensor([30.0, 40.0, 35.0]).view(1, 3, 1, 1)
    _check_candidates(head, True)
    # NaN logits kill a pixel; -inf everywhere gives score 0 / class 0
    head["logits"][0, 2, 3, 5:20] = float("nan")
    head["logits"][0, :, 4, :] = float("-inf")
    _check_candidates(head, True)
    # min_confidence <= 0: masked / out-of-partition candidates survive with score 0
    head = synth.make_head_outputs(1, 3, 4, 64, seed=6, n_objects=2)
    _check_candidates(head, True, min_conf=0.0)
    _check_candidates(head, False, min_conf=-1.0)
    # nothing passes
    head["logits"][:] = -20.0
    assert _check_candidates(head, True) == 0


def test_fused_candidates_multi_task():
    """Two tasks on one stride: categories offset by the task's class count (range_decoder.py:77)."""
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    h0 = synth.make_head_outputs(2, 3, 8, 128, seed=1, n_objects=4)
    h1 = synth.make_head_outputs(2, 2, 8, 128, seed=2, n_objects=4)
    ms = {1: {"cart": h0["cart"], "mask": h0["mask"], 0: {"logits": h0["logits"], "regressands": h0["regressands"]},
              1: {"logits": h1["logits"], "regressands": h1["regressands"]}}}
    tasks = {0: ["a", "b", "c"], 1: ["d", "e"]}
    ms_dev = {1: {"cart": h0["cart"].to(DEV), "mask": h0["mask"].to(DEV),
                  0: {k: v.to(DEV) for k, v in ms[1][0].items()}, 1: {k: v.to(DEV) for k, v in ms[1][1].items()}}}
    dec = RangeDecoder(True, True, *SBR)
    cand = dec.candidates(ms_dev, PP, tasks)
    got = unpack_candidates(cand, cand.count())
    params, scores, cats = oracle.range_decoder_decode(ms, PP, tasks, True, True, *SBR, return_candidates=True)
    live = scores.numpy() >= np.float32(0.1)
    assert set(zip(got["sweep"].tolist(), got["k"].tolist())) == set(zip(*[a.tolist() for a in np.nonzero(live)]))
    assert np.array_equal(got["category"], cats.numpy()[got["sweep"], got["k"]])
    assert got["category"].max() >= 3


def test_decode_fp16_inputs():
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    head = synth.make_head_outputs(1, 3, 8, 128, seed=8, n_objects=4)
    h16 = {k: (v.half() if v.dtype == torch.float32 else v) for k, v in head.items()}
    dec = RangeDecoder(True, True, *SBR)
    cand = dec.candidates(ms_outputs(to_dev(h16, DEV)), PP, {0: ["a", "b", "c"]})
    got = unpack_candidates(cand, cand.count())
    params, scores, cats = oracle.range_decoder_decode(ms_outputs(h16), PP, {0: ["a", "b", "c"]}, True, True, *SBR,
                                                       return_candidates=True)
    scores = scores.float().numpy()
    live = scores >= float(torch.tensor(0.1, dtype=torch.float16))
    ref_set = set(zip(*[a.tolist() for a in np.nonzero(live)]))
    got_set = set(zip(got["sweep"].tolist(), got["k"].tolist()))
    # fp16 sigmoid has ~1e-3 resolution: allow disagreement only right at the threshold
    assert all(abs(scores[b, k] - 0.1) < 2e-3 for b, k in ref_set ^ got_set)


# ------------------------------------------------------------------------------------------------------
# autocast: half-precision heads next to float32 cart (detector.py:329-333 runs decode under autocast;
# coding.py:126-128 widens regressands and cart separately, the result takes the regressands' dtype)
# ------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_decode_range_view_mixed_dtypes(dtype):
    from rv3d.math.ops.coding import decode_range_view
    head = synth.make_head_outputs(2, 3, 8, 256, 15, n_objects=6)
    reg, cart = head["regressands"].to(dtype), head["cart"]          # cart stays float32
    for flag in (True, False):
        ref = oracle.decode_range_view(reg, cart, flag)
        out = decode_range_view(reg.to(DEV), cart.to(DEV), flag)
        assert out.dtype == dtype and ref.dtype == dtype
        assert torch.equal(out.cpu(), ref)
    # and it is NOT what rounding cart to the heads' dtype first would give
    lossy = oracle.decode_range_view(reg, cart.to(dtype), True)
    assert not torch.equal(lossy, oracle.decode_range_view(reg, cart, True))


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_candidates_mixed_dtypes(dtype):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    head = synth.make_head_outputs(2, 3, 16, 256, seed=16, n_objects=8)
    hm = dict(head, logits=head["logits"].to(dtype), regressands=head["regressands"].to(dtype))
    tasks = {0: ["a", "b", "c"]}
    dec = RangeDecoder(True, True, *SBR)
    cand = dec.candidates(ms_outputs(to_dev(hm, DEV)), PP, tasks)
    got = unpack_candidates(cand, cand.count())
    params, scores, cats = oracle.range_decoder_decode(ms_outputs(hm), PP, tasks, True, True, *SBR, return_candidates=True)
    assert scores.dtype == dtype and params.dtype == dtype
    thr = float(torch.tensor(PP["min_confidence"], dtype=dtype))
    sc = scores.float().numpy()
    ref_set = set(zip(*[a.tolist() for a in np.nonzero(sc >= thr)]))
    got_set = set(zip(got["sweep"].tolist(), got["k"].tolist()))
    # a half-precision sigmoid has ~1e-3 (f16) / ~8e-3 (bf16) resolution: disagreement only right at the threshold
    band = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert all(abs(sc[b, k] - thr) < band for b, k in ref_set ^ got_set)
    assert len(ref_set & got_set) > 500
    both = sorted(ref_set & got_set)
    pos = {bk: i for i, bk in enumerate(zip(got["sweep"].tolist(), got["k"].tolist()))}
    rows = np.array([pos[bk] for bk in both])
    bi, ki = np.array([b for b, _ in both]), np.array([k for _, k in both])
    # boxes: one rounding of the same fp64 value to the half type on both sides
    assert np.array_equal(got["boxes"][rows, :7], params.float().numpy()[bi, ki])
    assert np.array_equal(got["category"][rows], cats.numpy()[bi, ki])


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16, torch.float32])
def test_partition_uses_cart_dtype(dtype):
    """cart.norm(dim=1) is rounded to cart's dtype before it meets the float32 bounds (range_decoder.py:140-143):
    pixels whose float32 norm is just above a bound can round down onto it and stay in the nearer partition."""
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    rng = np.random.default_rng(77)
    B, C, H, W = 1, 2, 8, 512
    n = B * H * W
    r = np.where(rng.random(n) < 0.5, 15.0, 30.0) + rng.uniform(-0.08, 0.08, size=n)
    az = rng.uniform(-math.pi, math.pi, size=n)
    z = rng.uniform(-1.0, 1.0, size=n)
    rho = np.sqrt(np.maximum(r * r - z * z, 0.0))
    cart = torch.from_numpy(np.stack([rho * np.cos(az), rho * np.sin(az), z], 0).reshape(1, 3, H, W).astype(np.float32)).to(dtype)
    head = {"logits": torch.from_numpy(rng.normal(2.0, 1.0, size=(B, C, H, W)).astype(np.float32)).to(dtype),
            "regressands": torch.from_numpy(rng.normal(0.0, 0.3, size=(B, 8, H, W)).astype(np.float32)).to(dtype),
            "cart": cart, "mask": torch.ones((B, 1, H, W), dtype=torch.bool)}
    tasks = {0: ["a", "b"]}
    dec = RangeDecoder(True, True, *SBR)
    cand = dec.candidates(ms_outputs(to_dev(head, DEV)), PP, tasks)
    got = unpack_candidates(cand, cand.count())
    _, scores, _ = oracle.range_decoder_decode(ms_outputs(head), PP, tasks, True, True, *SBR, return_candidates=True)
    thr = float(torch.tensor(PP["min_confidence"], dtype=dtype))
    ref_set = set(zip(*[a.tolist() for a in np.nonzero(scores.float().numpy() >= thr)]))
    got_set = set(zip(got["sweep"].tolist(), got["k"].tolist()))
    assert ref_set == got_set          # logits ~ N(2, 1): nothing sits at the score threshold
    if dtype != torch.float32:         # the test has teeth: float32 norms would assign some pixels differently
        d32 = cart.float().norm(dim=1)
        dt = cart.norm(dim=1).float()
        assert int(((d32 > 15) != (dt > 15)).sum() + ((d32 > 30) != (dt > 30)).sum()) > 0
