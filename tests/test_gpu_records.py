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
