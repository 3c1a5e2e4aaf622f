"""Detection wire format through the C ABI vs the oracle's column restatement: every field
bit-exact, decoder order preserved, range filter identical."""
import numpy as np
import pytest
import torch

from oracle import assign_oracle
from tests import synth
from tests.util import PP, SBR, ms_outputs, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _detections(seed=3, B=3):
    from rv3d.nn.decoders.range_decoder import RangeDecoder
    head = synth.make_head_outputs(B, 3, 16, 256, seed=seed, n_objects=10)
    dec = RangeDecoder(True, True, *SBR)
    return dec.decode(ms_outputs(to_dev(head, DEV)), dict(PP, nms_mode="HARD"), {0: ["a", "b", "c"]})


@pytest.mark.parametrize("max_range", [None, 30.0])
def test_build_records_vs_oracle(max_range):
    from rv3d.math.ops.coding import RECORD_DTYPE, build_records
    params, scores, cats, bidx = _detections()
    stamps = [315969904359876000 + 100_000_000 * b for b in range(3)]
    rec = build_records(params, scores, cats, bidx, stamps, max_range_m=max_range)
    ref = assign_oracle.detection_rows(params.cpu(), scores.cpu(), cats.cpu(), bidx.cpu(), stamps, max_range)
    assert rec.dtype == RECORD_DTYPE and len(rec) == len(ref["score"]) > 50
    if max_range is not None:
        assert len(rec) < params.shape[0] and rec["range_m"].max() <= max_range
    for k, v in ref.items():
        assert np.array_equal(rec[k], v), k
        assert rec[k].dtype == v.dtype, k


def test_build_dataframe_columns_and_joins():
    from rv3d.math.ops.coding import SCHEMA, build_dataframe
    params, scores, cats, bidx = _detections(seed=4)
    uuids = {"batch_index": [0, 2], "log_id": ["log-a", "log-c"], "timestamp_ns": [11, 33]}    # sweep 1 has no uuid row
    df = build_dataframe(params, scores, cats, bidx, uuids, ["CAR", "BUS", "PED"])
    assert list(df) == list(SCHEMA)
    keep = (bidx.cpu().numpy().astype(int) != 1).reshape(-1)
    assert len(df["score"]) == keep.sum() and set(df["batch_index"].tolist()) == {0, 2}
    assert np.array_equal(df["score"], scores.cpu().numpy().reshape(-1)[keep])               # order kept, inner join on batch_index
    assert np.array_equal(df["qw"], params.cpu().numpy()[keep, 6])
    names = np.array(["CAR", "BUS", "PED"], dtype=object)[cats.cpu().numpy().astype(int).reshape(-1)[keep]]
    assert np.array_equal(df["category"], names)
    assert np.array_equal(df["timestamp_ns"], np.where(df["batch_index"] == 0, 11, 33))
    assert set(df["log_id"].tolist()) == {"log-a", "log-c"}
    for k, dt in SCHEMA.items():
        assert df[k].dtype == dt, k


def test_build_records_empty():
    from rv3d.math.ops.coding import build_records
    e = torch.empty((0,), device=DEV)
    assert len(build_records(torch.empty((0, 10), device=DEV), e, e, e)) == 0


def test_prepare_for_evaluation_sort_unique_and_grouping():
    """detector.py:573-584 (range filter, sort by score descending, unique) and :366-380 (per-sweep groups) on the device
    record stream vs the oracle's restatement over numpy columns.  Duplicated detections (the same rows appended twice,
    plus rows that only differ in one field) exercise unique(); the survivors are compared as a SET (polars leaves their
    order unspecified) and their score order is checked."""
    from rv3d.math.ops.coding import build_records_device, group_by_sweep, prepare_for_evaluation
    params, scores, cats, bidx = _detections(seed=5)
    n = params.shape[0]
    dup = torch.arange(0, n, 3, device=DEV)
    near = params[dup].clone(); near[:, 3] += 1e-3                                   # same row but for one field: NOT a duplicate
    P = torch.cat([params, params[dup], near]); S = torch.cat([scores, scores[dup], scores[dup]])
    C = torch.cat([cats, cats[dup], cats[dup]]); Bi = torch.cat([bidx, bidx[dup], bidx[dup]])
    stamps = [315969904359876000 + 100_000_000 * b for b in range(3)]
    rec = prepare_for_evaluation(P, S, C, Bi, stamps, 40.0)
    names, want = assign_oracle.prepare_for_evaluation_rows(P.cpu(), S.cpu(), C.cpu(), Bi.cpu(), stamps, 40.0)
    got = {tuple(rec[k][i].item() for k in names) for i in range(len(rec))}
    assert len(rec) == len(got) == len(want) and got == set(want)
    assert len(rec) < P.shape[0] - dup.numel() + 1 and len(rec) > n // 2              # duplicates and far rows are gone
    assert np.all(np.diff(rec["score"]) <= 0)                                        # score descending
    # per-sweep groups of the unfiltered stream
    r, cnt = build_records_device(params, scores, cats, bidx, stamps)
    off = group_by_sweep(r, cnt, 3).cpu().numpy()
    b = bidx.cpu().numpy().astype(int).reshape(-1)
    assert off[0] == 0 and off[-1] == n and np.array_equal(off, np.searchsorted(b, np.arange(4)))
    e = torch.empty((0,), device=DEV)
    assert len(prepare_for_evaluation(torch.empty((0, 10), device=DEV), e, e, e, None, 10.0)) == 0
