"""oracle/assign_oracle.py against vectors minted from the verbatim torchbox3d.math.ops.assignment
(tests/golden/make_golden_assign.py)."""
import numpy as np
import pytest
import torch

from oracle import assign_oracle
from tests.conftest import GOLDEN

CFGS = {
    "bev": dict(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=5, normalize_affinities=False, sigma=1.0),
    "gauss": dict(affinity_fn="gaussian", enable_azimuth_invariant_targets=True, k=3, normalize_affinities=True, sigma=0.7),
    "gauss_raw": dict(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=False, k=100, normalize_affinities=False, sigma=1.5),
    # the production setting (conf/model/range_view.yaml:126 k = .inf, baseline.yaml:43-45 GAUSSIAN, sigma 0.75)
    "gauss_inf": dict(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=True, k=float("inf"), normalize_affinities=False, sigma=0.75),
    "gauss_k4": dict(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=True, k=4, normalize_affinities=False, sigma=0.75),
    "bev_k1": dict(affinity_fn="BEV", enable_azimuth_invariant_targets=False, k=1, normalize_affinities=False, sigma=1.0),
}


@pytest.mark.parametrize("tag", list(CFGS))
def test_compute_classification_targets_matches_reference(tag):
    g = np.load(GOLDEN / "assign.npz")
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    res = assign_oracle.compute_classification_targets(t("input"), t("target"), t("labels"), t("cart"), CFGS[tag], t("mask"),
                                                       t("panoptics"), 3)
    for name, r in zip(("affinities", "foreground", "background", "reg_weights"), res):
        assert np.array_equal(r.numpy(), g[f"{tag}_{name}"]), name
    assert g[f"{tag}_foreground"].sum() > 20 and g[f"{tag}_affinities"].max() > 0
