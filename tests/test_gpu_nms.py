"""CUDA suppression kernels (through the C ABI) vs the oracle and the golden vectors.
Bars: rotated IoU bit-exact; keep-sets bit-exact on identical (boxes, scores); weighted-merge rows
within 1e-5 relative (they are fp64 accumulations of float32 inputs)."""
import math

import numpy as np
import pytest
import torch

import oracle
from tests import synth
from tests.conftest import GOLDEN
from tests.util import PP, SBR, ms_outputs, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _rand_pairs(n, seed, overlap=True):
    rng = np.random.default_rng(seed)
    a = np.empty((n, 7), np.float32)
    a[:, :2] = rng.uniform(-50, 50, (n, 2)); a[:, 2] = rng.normal(0, 1, n)
    a[:, 3:6] = np.exp(rng.normal(0.8, 0.6, (n, 3))); a[:, 6] = rng.uniform(-2 * math.pi, 2 * math.pi, n)
    b = a.copy()
    b[:, :2] += rng.normal(0, 1.5 if overlap else 30, (n, 2)).astype(np.float32)
    b[:, 2] += rng.normal(0, 0.5, n).astype(np.float32)
    b[:, 3:6] *= np.exp(rng.normal(0, 0.2, (n, 3))).astype(np.float32)
    b[:, 6] += rng.normal(0, 0.4, n).astype(np.float32)
    return a, b


def test_rotated_iou_bit_exact_and_iou3d():
    from rv3d.math.ops.iou import iou_3d_axis_aligned
    a, b = _rand_pairs(200_000, 0)
    # special cases: identical, disjoint, 90-degree rotated squares, half shift, degenerate area
    sp_a = np.array([[0, 0, 0, 2, 2, 1, 0], [0, 0, 0, 2, 2, 1, 0], [0, 0, 0, 2, 2, 1, 0], [0, 0, 0, 2, 2, 1, 0],
                     [0, 0, 0, 1e-8, 1e-8, 1, 0], [5, 5, 0, 4, 2, 1, 0.3]], np.float32)
    sp_b = np.array([[0, 0, 0, 2, 2, 1, 0], [10, 0, 0, 2, 2, 1, 0], [0, 0, 0, 2, 2, 1, math.pi / 2], [1, 0, 0, 2, 2, 1, 0],
                     [0, 0, 0, 1, 1, 1, 0], [5, 5, 0, 4, 2, 1, 0.3 + math.pi]], np.float32)
    a, b = np.concatenate([sp_a, a]), np.concatenate([sp_b, b])
    i3, bev = iou_3d_axis_aligned(torch.from_numpy(a).to(DEV), torch.from_numpy(b).to(DEV))
    ref3, refbev = oracle.iou_3d_axis_aligned(torch.from_numpy(a), torch.from_numpy(b))
    bev, i3 = bev.cpu().numpy(), i3.cpu().numpy()
    assert np.array_equal(bev.view(np.uint32), refbev.numpy().view(np.uint32)), \
        f"{(bev != refbev.numpy()).sum()} of {len(bev)} BEV IoUs differ"
    np.testing.assert_allclose(i3, ref3.numpy(), rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(bev[:4], [1.0, 0.0, 1.0, 1 / 3], atol=1e-6)
    assert bev[4] == 0.0 and abs(bev[5] - 1.0) < 1e-5
    assert (bev > 0.3).mean() > 0.2          # the sample really exercises overlapping pairs


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (300, 2), (5000, 3), (20000, 4)])
def test_nms_rotated_keep_set(n, seed):
    from rv3d.math.ops.nms import nms_rotated
    cub, sc, _ = synth.make_nms_candidates(1, n, 1, max(n // 150, 1), seed)
    boxes = cub[0][:, [0, 1, 3, 4, 6]].clone()
    boxes[:, -1] = -boxes[:, -1].rad2deg()
    ref = oracle.nms_rotated(boxes, sc[0], torch.as_tensor(0.3))
    got = nms_rotated(boxes.to(DEV), sc[0].to(DEV), torch.as_tensor(0.3))
    assert got.dtype == torch.int64
    assert torch.equal(got.cpu(), ref), f"kept {len(got)} vs {len(ref)}"


def test_nms_rotated_ties_and_thresholds():
    from rv3d.math.ops.nms import nms_rotated
    cub, sc, _ = synth.make_nms_candidates(1, 3000, 1, 10, 9)
    boxes = cub[0][:, [0, 1, 3, 4, 6]].clone()
    boxes[:, -1] = -boxes[:, -1].rad2deg()
    scores = (sc[0] * 8).round() / 8          # massive score ties -> order = (score desc, index asc)
    for thr in (0.0, 0.1, 0.5, 0.9, 1.0):
        ref = oracle.nms_rotated(boxes, scores, torch.as_tensor(thr))
        got = nms_rotated(boxes.to(DEV), scores.to(DEV), thr)
        assert torch.equal(got.cpu(), ref), thr


@pytest.mark.parametrize("mode", ["hard", "weighted"])
def test_batched_nms_golden(mode):
    from rv3d.math.ops.nms import batched_multiclass_nms
    g = np.load(GOLDEN / f"nms_{mode}.npz")
    o = batched_multiclass_nms(torch.from_numpy(g["cuboids"]).to(DEV), torch.from_numpy(g["scores"]).to(DEV),
                               torch.from_numpy(g["categories"]).to(DEV), 500, 20, 0.3, 0.1, mode)
    assert np.array_equal(o[1].cpu().numpy(), g["out_scores"])
    assert np.array_equal(o[2].cpu().numpy(), g["out_categories"])
    assert np.array_equal(o[3].cpu().numpy(), g["out_batch_index"])
    if mode == "hard":
        assert np.array_equal(o[0].cpu().numpy(), g["out_cuboids"])          # kept rows are input rows
    else:
        np.testing.assert_allclose(o[0].cpu().numpy(), g["out_cuboids"], rtol=1e-5, atol=1e-5)
    assert o[2].dtype == torch.float32 and o[3].dtype == torch.float32


@pytest.mark.parametrize("mode", ["HARD", "WEIGHTED"])
# the last case has 24 x 26 = 624 (sweep, class) segments: above 512 the output offsets come from the scan kernel,
# below from the pack kernel itself
# (1, 40000, 1, 3000, 30000, 2200): more than 2048 kept boxes per segment -> the kept-box grid lives in global memory;
# in WEIGHTED mode the candidates behind the scan go to the grid-wide tail kernel in both forms, with the num_pre_nms
# cut inside the tail (cases 2 and 5)
@pytest.mark.parametrize("cfg", [(3, 20000, 5, 40, 50000, 1000), (2, 30000, 2, 25, 4000, 50), (1, 6000, 26, 30, 50000, 7),
                                 (24, 1500, 26, 20, 50000, 5), (1, 40000, 1, 3000, 30000, 2200)])
def test_batched_nms_vs_oracle(mode, cfg):
    from rv3d.math.ops.nms import batched_multiclass_nms
    B, K, C, M, pre, post = cfg
    cub, sc, ca = synth.make_nms_candidates(B, K, C, M, seed=K + C)
    ref = oracle.batched_multiclass_nms(cub, sc, ca, pre, post, 0.3, 0.1, mode)
    got = batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), pre, post, 0.3, 0.1, mode)
    assert got[1].shape == ref[1].shape
    assert torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[2].cpu(), ref[2]) and torch.equal(got[3].cpu(), ref[3])
    if mode == "HARD":
        assert torch.equal(got[0].cpu(), ref[0])
    else:
        np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-5)


def test_weighted_nms_wrapper():
    from rv3d.math.ops.nms import weighted_nms
    cub, sc, _ = synth.make_nms_candidates(1, 4000, 1, 20, 21)
    cu, s = cub[0], sc[0]
    boxes = torch.cat([cu[:, :2] - cu[:, 3:5] / 2, cu[:, :2] + cu[:, 3:5] / 2, cu[:, 6:7]], -1)
    d2m = torch.cat([cu[:, :6], cu[:, 6:7].sin(), cu[:, 6:7].cos()], 1)
    rk, ro, rc = oracle.weighted_nms(boxes, d2m, s, 0.3, 0.5)
    gk, go, gc = weighted_nms(boxes.to(DEV), d2m.to(DEV), s.to(DEV), 0.3, 0.5)
    assert torch.equal(gk.cpu(), rk) and torch.equal(gc.cpu(), rc)
    np.testing.assert_allclose(go.cpu().numpy(), ro.numpy(), rtol=1e-5, atol=1e-5)
    # merge_thresh below nms_thresh: a box can join several merge sets and still be kept later
    rk, ro, rc = oracle.weighted_nms(boxes, d2m, s, 0.6, 0.2)
    gk, go, gc = weighted_nms(boxes.to(DEV), d2m.to(DEV), s.to(DEV), 0.6, 0.2)
    assert torch.equal(gk.cpu(), rk) and torch.equal(gc.cpu(), rc)
    np.testing.assert_allclose(go.cpu().numpy(), ro.numpy(), rtol=1e-5, atol=1e-5)


def test_per_sweep_entry_points_and_errors():
    from rv3d.math.ops.nms import batched_multiclass_nms, hard_multiclass_nms, weighted_multiclass_nms
    cub, sc, ca = synth.make_nms_candidates(1, 3000, 3, 12, 33)
    for fn, ofn in ((hard_multiclass_nms, oracle.hard_multiclass_nms), (weighted_multiclass_nms, oracle.weighted_multiclass_nms)):
        ref = ofn(cub[0], sc[0], ca[0], 0.3, 1000, 25)
        got = fn(cub[0].to(DEV), sc[0].to(DEV), ca[0].to(DEV), 0.3, 1000, 25)
        assert torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[2].cpu(), ref[2])
        np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-5)
    with pytest.raises(NotImplementedError):
        batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), 10, 10, 0.3, 0.1, "soft")
    out = batched_multiclass_nms(cub.to(DEV), sc.to(DEV) * 0, ca.to(DEV), 10, 10, 0.3, 0.1, "hard")
    assert [tuple(o.shape) for o in out] == [(0, 7), (0, 1), (0, 1), (0, 1)]
    with pytest.raises(RuntimeError):
        batched_multiclass_nms(cub, sc, ca, 10, 10, 0.3, 0.1, "hard")           # CPU tensors: no fallback


@pytest.mark.parametrize("mode", ["hard", "weighted", "nonms"])
def test_range_decoder_pipeline_golden(mode):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    g = np.load(GOLDEN / f"pipeline_{mode}.npz")
    head = {k: torch.from_numpy(g[k]).to(DEV) for k in ("logits", "regressands", "cart", "mask")}
    pp = dict(PP, nms_mode="HARD" if mode == "nonms" else mode.upper())
    dec = RangeDecoder(True, mode != "nonms", *SBR)
    p, s, c, b = dec.decode(ms_outputs(head), pp, {0: ["A", "B", "C"]}, use_nms=mode != "nonms", data=None)
    assert tuple(p.shape) == g["params"].shape and s.shape[0] == g["scores"].shape[0]
    np.testing.assert_allclose(s.cpu().numpy(), g["scores"], rtol=1e-6)
    assert np.array_equal(c.cpu().numpy(), g["categories"]) and np.array_equal(b.cpu().numpy(), g["batch_index"])
    np.testing.assert_allclose(p.cpu().numpy(), g["params"], rtol=1e-5, atol=1e-5)
    assert c.dtype == (torch.int64 if mode == "nonms" else torch.float32)


@pytest.mark.parametrize("shape,mode", [((2, 3, 64, 2650), "HARD"), ((2, 3, 64, 2650), "WEIGHTED"), ((1, 26, 64, 1800), "HARD")])
def test_range_decoder_full_size_vs_oracle(shape, mode):
    """Full Waymo / AV2 shape, stage-wise on identical inputs: the oracle's NMS runs on the
    candidates the CUDA decode produced, so both see the same (boxes, scores)."""
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    from rv3d.math.ops.nms import batched_multiclass_nms
    B, C, H, W = shape
    head = synth.make_head_outputs(B, C, H, W, seed=17, n_objects=48, fp_rate=0.05, distinct_scores=False)
    pp = dict(PP, nms_mode=mode)
    tasks = {0: [f"c{i}" for i in range(C)]}
    dec = RangeDecoder(True, True, *SBR)
    p, s, c, b = dec.decode(ms_outputs(to_dev(head, DEV)), pp, tasks)
    # same candidates, densified, through the oracle
    from tests.util import unpack_candidates
    cand = dec.candidates(ms_outputs(to_dev(head, DEV)), pp, tasks)
    n = cand.count()
    u = unpack_candidates(cand, n)
    K = cand.total_candidates
    cub = torch.zeros(B, K, 7); sc = torch.zeros(B, K); ca = torch.zeros(B, K, dtype=torch.int64)
    cub[u["sweep"], u["k"]] = torch.from_numpy(u["boxes"]); sc[u["sweep"], u["k"]] = torch.from_numpy(u["score"])
    ca[u["sweep"], u["k"]] = torch.from_numpy(u["category"])
    ref = oracle.batched_multiclass_nms(cub, sc, ca, pp["num_pre_nms"], pp["num_post_nms"], 0.3, 0.1, mode)
    assert s.shape == ref[1].shape
    assert torch.equal(s.cpu(), ref[1]) and torch.equal(c.cpu(), ref[2]) and torch.equal(b.cpu(), ref[3])
    refp = torch.cat([ref[0][:, :-1], oracle.yaw_to_quat(ref[0][:, -1:])], -1)
    np.testing.assert_allclose(p.cpu().numpy(), refp.numpy(), rtol=1e-5, atol=1e-5)
    # size-independent properties: ordered (sweep asc, class asc, score desc); NMS is idempotent
    key = np.stack([b.cpu().numpy(), c.cpu().numpy(), -s.cpu().numpy()], 1)
    assert (np.lexsort(key.T[::-1]) == np.arange(len(key))).all()
    if mode == "HARD":
        kept7 = ref[0].to(DEV)
        again = batched_multiclass_nms(kept7[None], s[None], c[None].long() + 1000 * b[None].long(), 50000, 1000, 0.3, 0.1, "HARD")
        assert again[1].shape == s.shape


@pytest.mark.parametrize("mode", ["HARD", "WEIGHTED"])
def test_single_class_stress_queue_overflow(mode):
    """BASELINE config 3 in miniature: one class, 60 k heavily clustered candidates (hundreds of boxes per
    object), num_pre_nms truncation active.  Drives the work queues past their capacity, so the in-place
    (hard) and serial-fallback (weighted) overflow paths are part of the parity check."""
    from rv3d.math.ops.nms import batched_multiclass_nms
    cub, sc, ca = synth.make_nms_candidates(1, 60000, 1, 48, seed=77, spread=40.0, frac_clustered=0.97)
    ref = oracle.batched_multiclass_nms(cub, sc, ca, 50000, 300, 0.3, 0.03, mode)
    got = batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), 50000, 300, 0.3, 0.03, mode)
    assert torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[2].cpu(), ref[2]) and torch.equal(got[3].cpu(), ref[3])
    if mode == "HARD":
        assert torch.equal(got[0].cpu(), ref[0])
    else:
        np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-5)


def test_degenerate_boxes():
    """Tiny, huge, NaN and duplicated boxes: the pruning structures must not change the result."""
    from rv3d.math.ops.nms import batched_multiclass_nms
    cub, sc, ca = synth.make_nms_candidates(1, 4000, 2, 10, seed=5)
    cub[0, 10:40, 3:5] = 1e-4                      # degenerate extents (vertex-inside tolerance dominates)
    cub[0, 50:60, 3:5] = 300.0                     # huge boxes: oversize list
    cub[0, 70:75, 0] = 5e6                         # far away: clamped into a border cell
    cub[0, 80:90] = cub[0, 100:110]                # exact duplicates (different scores)
    cub[0, 120, 0] = float("nan")                  # NaN never interacts
    for mode in ("HARD", "WEIGHTED"):
        ref = oracle.batched_multiclass_nms(cub, sc, ca, 50000, 1000, 0.3, 0.1, mode)
        got = batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), 50000, 1000, 0.3, 0.1, mode)
        assert torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[2].cpu(), ref[2]), mode
        np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-5, equal_nan=True)
    # negative threshold: even disjoint boxes suppress each other -> one box per class survives
    ref = oracle.batched_multiclass_nms(cub, sc, ca, 50000, 1000, -1.0, 0.1, "HARD")
    got = batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), 50000, 1000, -1.0, 0.1, "HARD")
    assert torch.equal(got[1].cpu(), ref[1]) and got[1].shape[0] <= 3


def test_config3_weighted_stress_200k():
    """BASELINE config 3: one class, 200 000 candidates >= 0.1 clustered around 256 objects,
    num_pre_nms = 200 000, WEIGHTED (TorchEx is not installable here: compared against the oracle)."""
    from rv3d.math.ops.nms import batched_multiclass_nms
    cub, sc, ca = synth.make_nms_candidates(1, 200_000, 1, 256, seed=123, spread=75.0, frac_clustered=0.97)
    sc = 0.1 + 0.9 * sc                                           # every candidate passes min_confidence
    ref = oracle.batched_multiclass_nms(cub, sc, ca, 200_000, 1000, 0.3, 0.1, "WEIGHTED")
    got = batched_multiclass_nms(cub.to(DEV), sc.to(DEV), ca.to(DEV), 200_000, 1000, 0.3, 0.1, "WEIGHTED")
    assert torch.equal(got[1].cpu(), ref[1]) and torch.equal(got[2].cpu(), ref[2])
    np.testing.assert_allclose(got[0].cpu().numpy(), ref[0].numpy(), rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("fp_rate", [0.95, 0.01])
def test_bench_workload_properties_full_size(fp_rate):
    """BASELINE config 1 at its full size and the bench's candidate density (16 Waymo-shaped sweeps, ~830 k candidates):
    too large for the CPU oracle, so the result is checked through the properties that define greedy NMS, evaluated on
    the device with the library's own all-pairs IoU:
      order       rows sorted by (sweep asc, class asc, score desc), at most num_post_nms per (sweep, class)
      independent no two kept boxes of a segment overlap by more than the threshold
      maximal     in segments that did not hit num_post_nms, every dropped candidate overlaps a kept box of
                  higher score by more than the threshold
      idempotent  NMS of the kept boxes keeps all of them"""
    from rv3d.math.ops.assignment import box_iou_rotated
    from rv3d.math.ops.nms import batched_multiclass_nms
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    from tests.util import unpack_candidates
    B, C, H, W = 16, 3, 64, 2650
    head = synth.make_head_outputs(B, C, H, W, seed=1000, n_objects=96, fp_rate=fp_rate)   # 0.95: the bench's density; 0.01: segments stay below num_post_nms
    pp = dict(PP, nms_mode="HARD")
    tasks = {0: ["a", "b", "c"]}
    dec = RangeDecoder(True, True, *SBR)
    ms = ms_outputs(to_dev(head, DEV))
    p, s, c, b = dec.decode(ms, pp, tasks, use_nms=True)
    assert p.shape[0] > (30000 if fp_rate > 0.5 else 2000)
    # yaw from the quaternion (qw, 0, 0, qz): the decoder's rows carry params(10)
    cand = dec.candidates(ms, pp, tasks)
    u = unpack_candidates(cand, cand.count())
    assert len(u["score"]) > (500_000 if fp_rate > 0.5 else 20_000)
    sc, cc, bc = s.cpu().numpy(), c.cpu().numpy().astype(int), b.cpu().numpy().astype(int)
    key = np.stack([bc, cc, -sc], 1)
    assert (np.lexsort(key.T[::-1]) == np.arange(len(key))).all()
    seg_out = bc * C + cc
    counts = np.bincount(seg_out, minlength=B * C)
    assert counts.max() <= pp["num_post_nms"]
    # kept boxes as (x, y, l, w, angle): detectron2 gets angle = -rad2deg(yaw) (nms.py:40); the radian routine checked with
    # here is mmcv's, whose rotation direction is the opposite one, so it sees +yaw for the same geometry.  The float32
    # degree round trip differs by ~1e-7 rad and the two flavours enumerate the vertices differently, so comparisons
    # carry a 1e-4 IoU margin.
    cand_boxes = torch.from_numpy(u["boxes"]).to(DEV)
    cand5 = torch.stack([cand_boxes[:, 0], cand_boxes[:, 1], cand_boxes[:, 3], cand_boxes[:, 4], cand_boxes[:, 6]], 1)
    cand_seg = u["sweep"] * C + u["category"]
    thr = 0.3
    # map kept rows back to candidates through (segment, score, position): scores are distinct per sweep in this generator
    checked = 0
    for seg in (0, 7, 23, 47):
        rows = np.nonzero(seg_out == seg)[0]
        idx = np.nonzero(cand_seg == seg)[0]
        cs = u["score"][idx]
        order = np.argsort(-cs, kind="stable")
        idx, cs = idx[order], cs[order]
        pos = np.searchsorted(-cs, -sc[rows])                                   # kept scores among the candidates'
        assert np.array_equal(cs[pos], sc[rows])
        np.testing.assert_allclose(u["boxes"][idx[pos], :3], p[rows, :3].cpu().numpy(), rtol=0, atol=0)   # same boxes
        kept5 = cand5[torch.from_numpy(idx[pos]).to(DEV)]
        iou_kk = box_iou_rotated(kept5, kept5)
        iou_kk.fill_diagonal_(0)
        assert float(iou_kk.max()) <= thr + 1e-4                                 # independent
        if len(rows) < pp["num_post_nms"]:
            kept_mask = np.zeros(len(idx), dtype=bool); kept_mask[pos] = True
            dropped = np.nonzero(~kept_mask)[0]
            iou_dk = box_iou_rotated(cand5[torch.from_numpy(idx[dropped]).to(DEV)], kept5)      # (dropped, kept)
            higher = torch.from_numpy(pos[None, :] < dropped[:, None]).to(DEV)                   # kept box ranks above the dropped one
            assert bool(((iou_dk > thr - 1e-4) & higher).any(dim=1).all())       # maximal
            checked += 1
    assert checked > 0 or fp_rate > 0.5
    # idempotent: feed every kept box back, one "sweep", class id = segment
    k7 = torch.cat([p[:, :6], 2.0 * torch.atan2(p[:, 9], p[:, 6])[:, None]], 1)
    again = batched_multiclass_nms(k7[None], s[None], torch.from_numpy(seg_out).to(DEV)[None], 50000, 1000, thr, 0.1, "HARD")
    assert again[1].shape[0] == s.shape[0]
