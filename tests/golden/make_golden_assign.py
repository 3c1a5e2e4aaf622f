"""Golden vectors for the training-time callers: run ONCE in the build container.

    python tests/golden/make_golden_assign.py        ->  tests/golden/assign.npz

``torchbox3d.math.ops.assignment`` is imported VERBATIM (compute_classification_targets, iou_2d/3d_axis_aligned,
_gaussian); its one native dependency, mmcv.ops.box_iou_rotated (not installed), is replaced by the oracle's C
routine -> the file pins the CONTROL FLOW (decode x2, per-instance masked_select / topk / masked_scatter_, masks),
not mmcv's IoU arithmetic, which stays "parity unpinned"."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.golden import make_golden as mg  # noqa: E402,F401  (import stubs + reference on sys.path)
from tests import synth  # noqa: E402
from oracle import assign_oracle  # noqa: E402

from torchbox3d.math.ops import assignment as ref  # noqa: E402

OUT = Path(__file__).resolve().parent


class Cfg(dict):
    __getattr__ = dict.__getitem__


def main():
    ref.box_iou_rotated = lambda a, b, aligned=False: assign_oracle.box_iou_rotated(a, b, aligned)
    out = {}
    d = synth.make_assignment_inputs(2, 3, 16, 128, seed=51)
    out.update({k: v.numpy() for k, v in d.items()})
    for tag, cfg in {
        "bev": Cfg(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=5, normalize_affinities=False, sigma=1.0),
        "gauss": Cfg(affinity_fn="gaussian", enable_azimuth_invariant_targets=True, k=3, normalize_affinities=True, sigma=0.7),
        "gauss_raw": Cfg(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=False, k=100, normalize_affinities=False, sigma=1.5),
        # the production setting (conf/model/range_view.yaml:126 k = .inf, baseline.yaml:43-45 GAUSSIAN, sigma 0.75)
        "gauss_inf": Cfg(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=True, k=float("inf"), normalize_affinities=False, sigma=0.75),
        "gauss_k4": Cfg(affinity_fn="GAUSSIAN", enable_azimuth_invariant_targets=True, k=4, normalize_affinities=False, sigma=0.75),
        "bev_k1": Cfg(affinity_fn="BEV", enable_azimuth_invariant_targets=False, k=1, normalize_affinities=False, sigma=1.0),
    }.items():
        res = ref.compute_classification_targets(d["input"].clone(), d["target"].clone(), d["labels"], d["cart"], cfg, d["mask"],
                                                 d["panoptics"], 3)
        for name, t in zip(("affinities", "foreground", "background", "reg_weights"), res):
            out[f"{tag}_{name}"] = t.numpy()
    # the two free-standing affinities on aligned pairs
    rng = np.random.default_rng(8)
    a = synth.make_nms_candidates(1, 600, 1, 40, seed=12)[0][0].clone()
    b = a + torch.from_numpy(rng.normal(0, 0.3, size=a.shape).astype(np.float32))
    out.update(pair_a=a.numpy(), pair_b=b.numpy(),
               pair_iou3d=ref.iou_3d_axis_aligned(a, b, normalize_affinities=False).numpy(),
               pair_iou3d_norm=ref.iou_3d_axis_aligned(a, b, normalize_affinities=True).numpy(),
               pair_iou2d=ref.iou_2d_axis_aligned(a, b, normalize_affinities=False).numpy())
    np.savez_compressed(OUT / "assign.npz", **out)
    print("written", OUT / "assign.npz", {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
