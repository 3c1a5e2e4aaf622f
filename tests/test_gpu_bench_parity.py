"""Parity AT THE JUDGED DENSITY: the exact inputs bench.py times (Waymo shape, seed 1000, fp_rate 0.95, every segment
hits num_post_nms) through the CUDA path vs the CPU oracle, keep-set for keep-set.  Also: results are owned by the caller
(a later call does not overwrite them) and the captured CUDA graph of the step reproduces the eager result."""
import numpy as np
import pytest
import torch

import oracle
from tests import synth
from tests.util import PP, SBR, ms_outputs, to_dev, unpack_candidates

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
B_BENCH, C, H, W = 16, 3, 64, 2650


@pytest.fixture(scope="module")
def bench_head():
    # bench.py make_inputs("waymo", 16, 1000, 0.95): the generator's stream depends on the batch size, so the whole
    # batch is generated and the first sweeps are used
    return synth.make_head_outputs(B_BENCH, C, H, W, seed=1000, n_objects=96, fp_rate=0.95, distinct_scores=False)


def _densify(cand, n, B):
    u = unpack_candidates(cand, n)
    K = cand.total_candidates
    cub = torch.zeros(B, K, 7); sc = torch.zeros(B, K); ca = torch.zeros(B, K, dtype=torch.int64)
    cub[u["sweep"], u["k"]] = torch.from_numpy(u["boxes"]); sc[u["sweep"], u["k"]] = torch.from_numpy(u["score"])
    ca[u["sweep"], u["k"]] = torch.from_numpy(u["category"])
    return cub, sc, ca, u


@pytest.mark.parametrize("mode", ["HARD", "WEIGHTED"])
def test_bench_inputs_keep_set_equals_oracle(bench_head, mode):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    nb = 2
    head = {k: v[:nb].contiguous() for k, v in bench_head.items()}
    pp = dict(PP, nms_mode=mode)
    tasks = {0: ["a", "b", "c"]}
    dec = RangeDecoder(True, True, *SBR)
    ms = ms_outputs(to_dev(head, DEV))
    p, s, c, b = dec.decode(ms, pp, tasks)
    cand = dec.candidates(ms, pp, tasks)
    n = cand.count()
    assert n > nb * 40_000                       # the judged density: ~52 k candidates per sweep
    cub, sc, ca, _ = _densify(cand, n, nb)
    ref = oracle.batched_multiclass_nms(cub, sc, ca, pp["num_pre_nms"], pp["num_post_nms"], 0.3, 0.1, mode)
    assert ref[1].shape[0] == nb * C * pp["num_post_nms"]          # every segment hits num_post_nms, as in the bench
    assert s.shape == ref[1].shape
    assert torch.equal(s.cpu(), ref[1]) and torch.equal(c.cpu(), ref[2]) and torch.equal(b.cpu(), ref[3])
    refp = torch.cat([ref[0][:, :-1], oracle.yaw_to_quat(ref[0][:, -1:])], -1)
    if mode == "HARD":
        assert torch.equal(p[:, :6].cpu(), refp[:, :6])            # kept rows are input rows
    np.testing.assert_allclose(p.cpu().numpy(), refp.numpy(), rtol=1e-5, atol=1e-5)


def test_results_are_owned_and_graph_replay_matches(bench_head):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    pp = dict(PP, nms_mode="HARD")
    tasks = {0: ["a", "b", "c"]}
    dec = RangeDecoder(True, True, *SBR)
    h0 = to_dev({k: v[:2].contiguous() for k, v in bench_head.items()}, DEV)
    h1 = to_dev({k: v[2:4].contiguous() for k, v in bench_head.items()}, DEV)
    first = dec.decode(ms_outputs(h0), pp, tasks)
    keep = [t.clone() for t in first]
    second = dec.decode(ms_outputs(h1), pp, tasks)                  # same shapes, same workspace
    assert not torch.equal(second[0], keep[0])
    for a, k in zip(first, keep):
        assert torch.equal(a, k), "a later decode() overwrote an earlier result"
    # the step has no host read: capture it once, replay it, same detections
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.graph(g, stream=side):
        det = dec.decode_async(ms_outputs(h0), pp, tasks)
    torch.cuda.current_stream().wait_stream(side)
    for _ in range(3):
        g.replay()
    got = det.result()
    for a, k in zip(got, keep):
        assert torch.equal(a, k)
    # new data in the captured input tensors -> the replay follows
    for k in h0:
        h0[k].copy_(h1[k])
    g.replay()
    got = det.result()
    for a, k in zip(got, second):
        assert torch.equal(a, k)
