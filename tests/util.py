"""Helpers shared by the GPU parity tests."""
from __future__ import annotations

import math

import numpy as np
import torch


def to_dev(head, dev="cuda:0"):
    return {k: v.to(dev) for k, v in head.items()}


def ms_outputs(head, stride=1, task=0):
    return {stride: {"cart": head["cart"], "mask": head["mask"],
                     task: {"logits": head["logits"], "regressands": head["regressands"]}}}


PP = {"num_pre_nms": 50000, "num_post_nms": 1000, "nms_threshold": 0.3, "min_confidence": 0.1, "nms_mode": "HARD"}
SBR = ([0, 15, 30], [15, 30, math.inf], [8, 2, 1])


def unpack_candidates(cand, n):
    """Candidates (device) -> dict of numpy arrays keyed by (sweep, candidate index)."""
    keys = cand.keys[:n].cpu().numpy().view(np.uint64)
    boxes = cand.boxes[:n].cpu().numpy()
    idx_bits = max(int(math.ceil(math.log2(cand.total_candidates))), 0) if cand.total_candidates > 1 else 0
    while (1 << idx_bits) < cand.total_candidates:
        idx_bits += 1
    seg = (keys >> np.uint64(cand.score_bits + idx_bits)).astype(np.int64)
    k = (keys & np.uint64((1 << idx_bits) - 1)).astype(np.int64)
    return {"sweep": seg // cand.total_classes, "category": seg % cand.total_classes, "k": k,
            "boxes": boxes[:, :7], "score": boxes[:, 7]}


def sort_rows(*cols):
    order = np.lexsort(tuple(reversed(cols)))
    return order
