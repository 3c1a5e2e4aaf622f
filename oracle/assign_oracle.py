"""CPU restatement of the training-time callers.  TEST INFRASTRUCTURE ONLY.

math/ops/assignment.py (paths relative to /root/reference) with mmcv's box_iou_rotated replaced by the oracle's C
routine (orc_rot_iou_aligned_mmcv: radians, mmcv's rotation direction, pinned by mmcv's published unit-test vector
to 1e-4; un-vendored, so its float32 bits stay unpinned), the control
flow is pinned by tests/golden/assign.npz (verbatim reference module, same substitution)."""
from __future__ import annotations

import numpy as np
import torch

from .rv_oracle import decode_range_view, mmcv_iou_pairs

XYLWA = [0, 1, 3, 4, 6]


def box_iou_rotated(a: torch.Tensor, b: torch.Tensor, aligned: bool = False) -> torch.Tensor:
    a, b = a.float().numpy(), b.float().numpy()
    if aligned:
        return torch.from_numpy(mmcv_iou_pairs(a, b))
    n, m = len(a), len(b)
    if n == 0 or m == 0:
        return torch.zeros((n, m))
    return torch.from_numpy(mmcv_iou_pairs(np.repeat(a, m, axis=0), np.tile(b, (n, 1))).reshape(n, m))


def iou_2d_axis_aligned(a, b, **kw):                                   # :64-73
    iou = box_iou_rotated(a[:, XYLWA].contiguous(), b[:, XYLWA].contiguous(), aligned=True).clamp(0.0, 1.0)
    if kw["normalize_affinities"]:
        raise UnboundLocalError("object_ious")
    return iou


def gaussian(a, b, **kw):                                              # :151-159
    d = torch.linalg.norm(a[:, :3] - b[:, :3], dim=-1)
    if kw["normalize_affinities"]:
        d = d - d.min()
    return torch.exp(-d / kw["sigma"] ** 2)


def compute_classification_targets(inp, target, labels, cart, cfg, mask, panoptics, background_index):
    """:76-148, instance by instance."""
    onehot = torch.nn.functional.one_hot(labels, background_index + 1).permute(0, 3, 1, 2)[:, :-1].float()
    fn = {"BEV": iou_2d_axis_aligned, "GAUSSIAN": gaussian}[str(cfg["affinity_fn"]).upper()]
    pds = decode_range_view(inp.detach(), cart, True)
    gts = decode_range_view(target, cart, bool(cfg["enable_azimuth_invariant_targets"]))
    aff = torch.zeros_like(target[:, 0:1])
    fgm = torch.zeros_like(target[:, 0:1])
    B = target.shape[0]
    ids_all = panoptics.reshape(B, *target.shape[2:])
    for i in range(B):
        for inst in torch.unique(ids_all[i]).tolist():
            if inst == 0:
                continue
            m = ids_all[i] == inst
            d_i = pds[i][:, m].t()
            g_i = gts[i][:, m].t()
            a_i = fn(d_i, g_i, **cfg)
            k = min(cfg["k"], len(a_i))                          # :129 (k may be .inf: conf/model/range_view.yaml:126)
            val, idx = a_i.topk(k)
            like = torch.zeros_like(a_i).scatter(0, idx, val)
            aff[i, 0][m] = like.type_as(aff)
            fgm[i, 0][m] = like.bool().type_as(aff)
    bgm = torch.logical_and(fgm.logical_not(), mask)
    return aff * onehot, fgm, bgm, onehot.any(dim=1, keepdim=True)


# --------------------------------------------------------------------------------------------
# detection wire format: math/ops/coding.py:31-76, nn/arch/detector.py:45-60,573-581.
# polars is not installed, so build_dataframe's joins are restated over numpy columns: parity unpinned for the
# frame plumbing; there is no arithmetic besides the float32 range norm.
# --------------------------------------------------------------------------------------------
def detection_rows(params, scores, categories, batch_index, timestamps_ns=None, max_range_m=None):
    """-> dict of numpy columns, decoder order kept, optionally range-filtered (detector.py:573-581)."""
    p = params.float().numpy()
    cols = {n: p[:, i].copy() for i, n in enumerate(("tx_m", "ty_m", "tz_m", "length_m", "width_m", "height_m", "qw", "qx", "qy", "qz"))}
    cols["score"] = scores.float().flatten().numpy()
    cols["category_index"] = categories.int().flatten().numpy()                  # coding.py:54
    cols["batch_index"] = batch_index.int().flatten().numpy()                    # coding.py:55
    if timestamps_ns is not None:
        ts = np.asarray(timestamps_ns, dtype=np.int64)
        cols["timestamp_ns"] = ts[cols["batch_index"]]
    norms = np.linalg.norm(p[:, :3], axis=-1)                                     # float32 in, float32 out (detector.py:573-576)
    cols["range_m"] = norms
    if max_range_m is not None:
        keep = norms.astype(np.float64) <= float(max_range_m)                     # .le(lit(cfg.max_range_m))
        cols = {k: v[keep] for k, v in cols.items()}
    return cols


def prepare_for_evaluation_rows(params, scores, categories, batch_index, timestamps_ns, max_range_m):
    """detector.py:573-584 on the detection frame: range filter, `.sort(score, descending=True)`, `.unique()` -> the set
    of distinct rows as a list of tuples sorted by score descending (polars leaves the order inside equal scores, and
    after unique(), unspecified: compare as sets + the score order)."""
    cols = detection_rows(params, scores, categories, batch_index, timestamps_ns, max_range_m)
    names = list(cols)
    rows = {tuple(cols[k][i].item() for k in names) for i in range(len(cols["score"]))}
    return names, sorted(rows, key=lambda r: -r[names.index("score")])

