"""The oracle's third-party restatements (detectron2/mmcv rotated IoU, TorchEx iou_bev, weighted NMS) vs
independent geometry: an fp64 Sutherland-Hodgman clip, OpenCV's rotatedRectangleIntersection, known
answers, and a brute-force Python restatement of the weighted-NMS rule.  CPU only.

These kernels are "parity unpinned" (their sources are not in the reference tree); this file is what
anchors them."""
import math

import numpy as np
import pytest
import torch

import oracle


def _corners(xc, yc, w, h, ang):
    """Corners of a w x h rectangle whose w-axis points along (cos ang, sin ang), counter-clockwise."""
    c, s = math.cos(ang), math.sin(ang)
    pts = []
    for sx, sy in ((0.5, 0.5), (-0.5, 0.5), (-0.5, -0.5), (0.5, -0.5)):
        pts.append((xc + c * sx * w - s * sy * h, yc + s * sx * w + c * sy * h))
    return pts


def _clip(subject, clipper):
    def inside(p, a, b):
        return (b[0] - a[0]) * (p[1] - a[1]) - (b[1] - a[1]) * (p[0] - a[0]) >= 0

    def isect(p, q, a, b):
        x1, y1, x2, y2, x3, y3, x4, y4 = *p, *q, *a, *b
        den = (x1 - x2) * (y3 - y4) - (y1 - y2) * (x3 - x4)
        t = ((x1 - x3) * (y3 - y4) - (y1 - y3) * (x3 - x4)) / den
        return (x1 + t * (x2 - x1), y1 + t * (y2 - y1))

    out = subject
    for i in range(len(clipper)):
        a, b = clipper[i], clipper[(i + 1) % len(clipper)]
        inp, out = out, []
        for j in range(len(inp)):
            p, q = inp[j], inp[(j + 1) % len(inp)]
            if inside(q, a, b):
                if not inside(p, a, b):
                    out.append(isect(p, q, a, b))
                out.append(q)
            elif inside(p, a, b):
                out.append(isect(p, q, a, b))
        if not out:
            return []
    return out


def _area(poly):
    return 0.5 * abs(sum(poly[i][0] * poly[(i + 1) % len(poly)][1] - poly[(i + 1) % len(poly)][0] * poly[i][1]
                         for i in range(len(poly))))


def _iou64(a, b):
    pa, pb = _corners(*a), _corners(*b)
    inter = _clip(pa, pb)
    ia = _area(inter) if len(inter) >= 3 else 0.0
    return ia / (a[2] * a[3] + b[2] * b[3] - ia)


def _random_boxes(n, seed):
    rng = np.random.default_rng(seed)
    a = np.stack([rng.uniform(-20, 20, n), rng.uniform(-20, 20, n), np.exp(rng.normal(1.0, 0.5, n)),
                  np.exp(rng.normal(0.5, 0.5, n)), rng.uniform(-math.pi, math.pi, n)], 1)
    b = a.copy()
    b[:, :2] += rng.normal(0, 1.0, (n, 2)); b[:, 2:4] *= np.exp(rng.normal(0, 0.2, (n, 2))); b[:, 4] += rng.normal(0, 0.5, n)
    return a.astype(np.float32), b.astype(np.float32)


def test_rot_iou_matches_fp64_clip_radians():
    """mmcv convention: angle in radians, w-axis along (cos a, -sin a)  => equals a CCW box at -a."""
    a, b = _random_boxes(3000, 0)
    got = oracle.rot_iou_pairs(a, b, 1.0)
    ref = np.array([_iou64((x[0], x[1], x[2], x[3], -x[4]), (y[0], y[1], y[2], y[3], -y[4])) for x, y in zip(a.astype(float), b.astype(float))])
    assert np.abs(got - ref).max() < 2e-4
    assert (ref > 0.3).mean() > 0.3


def test_rot_iou_degrees_matches_negated_yaw():
    """nms.py:40 passes -rad2deg(yaw): with detectron2's clockwise-positive degrees that is a box whose
    length points along (cos yaw, sin yaw)."""
    a, b = _random_boxes(1000, 1)
    ad, bd = a.copy(), b.copy()
    ad[:, 4] = -np.rad2deg(a[:, 4]); bd[:, 4] = -np.rad2deg(b[:, 4])
    got = oracle.rot_iou_pairs(ad, bd, 0.01745329251)
    ref = np.array([_iou64(tuple(x), tuple(y)) for x, y in zip(a.astype(float), b.astype(float))])
    assert np.abs(got - ref).max() < 2e-4


def test_rot_iou_vs_opencv():
    cv2 = pytest.importorskip("cv2")
    a, b = _random_boxes(500, 2)
    got = oracle.rot_iou_pairs(a, b, 1.0)
    for k in range(len(a)):
        ra = ((float(a[k, 0]), float(a[k, 1])), (float(a[k, 2]), float(a[k, 3])), -math.degrees(float(a[k, 4])))
        rb = ((float(b[k, 0]), float(b[k, 1])), (float(b[k, 2]), float(b[k, 3])), -math.degrees(float(b[k, 4])))
        kind, pts = cv2.rotatedRectangleIntersection(ra, rb)
        inter = cv2.contourArea(cv2.convexHull(pts)) if kind != 0 and pts is not None and len(pts) >= 3 else 0.0
        ref = inter / (a[k, 2] * a[k, 3] + b[k, 2] * b[k, 3] - inter)
        assert abs(got[k] - ref) < 5e-4, (k, got[k], ref)


def test_rot_iou_known_answers():
    def iou(a, b, scale=1.0):
        return float(oracle.rot_iou_pairs(np.array([a], np.float32), np.array([b], np.float32), scale)[0])

    assert iou([0, 0, 2, 2, 0], [0, 0, 2, 2, 0]) == pytest.approx(1.0, abs=1e-6)          # identical
    assert iou([0, 0, 2, 2, 0], [10, 0, 2, 2, 0]) == 0.0                                 # disjoint
    assert iou([0, 0, 2, 2, 0], [0, 0, 2, 2, math.pi / 2]) == pytest.approx(1.0, abs=1e-6)  # 90 deg square
    assert iou([0, 0, 2, 2, 0], [1, 0, 2, 2, 0]) == pytest.approx(1 / 3, abs=1e-6)        # half shift
    assert iou([0, 0, 1e-8, 1e-8, 0], [0, 0, 1, 1, 0]) == 0.0                            # area < 1e-14
    assert iou([0, 0, 4, 2, 90], [0, 0, 2, 4, 0], 0.01745329251) == pytest.approx(1.0, abs=1e-6)   # degrees


def test_iou_bev_matches_fp64_clip():
    """(x1,y1,x2,y2,ry): axis-aligned extents rotated counter-clockwise by +ry about the centre."""
    a, b = _random_boxes(3000, 3)

    def to_bev(x):
        return np.stack([x[:, 0] - x[:, 2] / 2, x[:, 1] - x[:, 3] / 2, x[:, 0] + x[:, 2] / 2, x[:, 1] + x[:, 3] / 2, x[:, 4]], 1)

    got = oracle.iou_bev_pairs(to_bev(a), to_bev(b))
    ref = np.array([_iou64(tuple(x), tuple(y)) for x, y in zip(a.astype(float), b.astype(float))])
    assert np.abs(got - ref).max() < 2e-4


def test_nms_rotated_is_sequential_greedy():
    rng = np.random.default_rng(4)
    a, _ = _random_boxes(400, 5)
    a[:, :2] = rng.uniform(-6, 6, (400, 2))
    scores = rng.permutation(400).astype(np.float32) / 400
    deg = a.copy(); deg[:, 4] = -np.rad2deg(a[:, 4])
    keep = oracle.nms_rotated(torch.from_numpy(deg), torch.from_numpy(scores), torch.as_tensor(0.3)).numpy()
    order = np.argsort(-scores, kind="stable")
    ref = []
    for i in order:
        if all(_iou64(tuple(a[k].astype(float)), tuple(a[i].astype(float))) <= 0.3 + 1e-4 for k in ref):
            ref.append(i)
    # the two only differ where an IoU sits within 1e-4 of the threshold
    assert len(set(keep) ^ set(ref)) <= 2
    assert (np.diff(scores[keep]) <= 0).all()


def test_weighted_nms_rule():
    """merge set of kept k = {k} + {j > k alive when k is kept with iou > merge_thresh};
    merged row = score-weighted mean, score column keeps s_k; rows beyond num_out stay zero."""
    rng = np.random.default_rng(6)
    a, _ = _random_boxes(300, 7)
    a[:, :2] = rng.uniform(-5, 5, (300, 2))
    scores = torch.from_numpy(rng.permutation(300).astype(np.float32) / 300 + 0.01)
    boxes = torch.from_numpy(np.stack([a[:, 0] - a[:, 2] / 2, a[:, 1] - a[:, 3] / 2, a[:, 0] + a[:, 2] / 2,
                                       a[:, 1] + a[:, 3] / 2, a[:, 4]], 1))
    data = torch.from_numpy(rng.normal(size=(300, 8)).astype(np.float32))
    keep, out, count = oracle.weighted_nms(boxes, data, scores, 0.3, 0.5)
    order = np.argsort(-scores.numpy(), kind="stable")
    alive = np.ones(300, bool)
    exp_keep, exp_rows, exp_cnt = [], [], []
    bnp = boxes.numpy()
    for pos, i in enumerate(order):
        if not alive[i]:
            continue
        members = [i]
        for j in order[pos + 1:]:
            if not alive[j]:
                continue
            iou = float(oracle.iou_bev_pairs(bnp[i:i + 1], bnp[j:j + 1])[0])
            if iou > np.float32(0.5):
                members.append(j)
            if iou > np.float32(0.3):
                alive[j] = False
        w = scores.numpy()[members].astype(np.float64)
        row = (data.numpy()[members].astype(np.float64) * w[:, None]).sum(0) / w.sum()
        exp_keep.append(i); exp_rows.append(np.append(row, scores.numpy()[i])); exp_cnt.append(len(members))
    assert keep.tolist() == exp_keep and count.tolist() == exp_cnt
    np.testing.assert_allclose(out.numpy(), np.array(exp_rows), rtol=1e-6, atol=1e-6)
