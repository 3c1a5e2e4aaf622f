"""The NMS kernels decide most (candidate, kept) pairs with a float32 approximate IoU and send only the pairs inside a
band around the threshold to the bit-exact routine.  This test measures that shortcut instead of trusting it: over
> 1e7 pairs, adversarial ones included (IoU within +-5 % of 0.3 / 0.5, aspect ratios up to 1:50, extents 0.05 - 100 m,
near-parallel and near-perpendicular edges, far-away centres), EVERY pair the shortcut decides must agree with the exact
routine, for both routines (detectron2-style rot_iou, mmdet3d-style iou_bev).  It prints the smallest distance between
an approximately-decided pair's exact IoU and the threshold (the safety margin actually observed)."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _decisions(a5, b5, thr, routine):
    from rv3d import _native as N
    from rv3d._util import ptr, stream_ptr
    n = a5.shape[0]
    dec = torch.empty(n, dtype=torch.int8, device=DEV)
    ap = torch.empty(n, dtype=torch.float32, device=DEV)
    ex = torch.empty(n, dtype=torch.float32, device=DEV)
    N.check(N.lib().rv3d_pair_decisions(ptr(a5.contiguous()), ptr(b5.contiguous()), n, float(thr), routine, ptr(dec), ptr(ap), ptr(ex),
                                        stream_ptr(torch.device(DEV))), "rv3d_pair_decisions")
    return dec, ap, ex


def _pairs(n, gen, kind, thr):
    """-> (xc, yc, l, w, yaw) x 2 in metres / radians."""
    u = lambda lo, hi: torch.empty(n, device=DEV).uniform_(lo, hi, generator=gen)   # noqa: E731
    nrm = lambda s: torch.empty(n, device=DEV).normal_(0, s, generator=gen)         # noqa: E731
    if kind == "generic":
        l, w = torch.exp(nrm(0.6) + 1.0), torch.exp(nrm(0.5) + 0.4)
    elif kind == "thin":                       # aspect ratios up to 1:50, small to large extents
        w = torch.exp(u(math.log(0.05), math.log(2.0)))
        l = w * torch.exp(u(0.0, math.log(50.0)))
    else:                                      # "extent": 0.05 m .. 100 m
        l, w = torch.exp(u(math.log(0.05), math.log(100.0))), torch.exp(u(math.log(0.05), math.log(100.0)))
        big, small = torch.maximum(l, w), torch.minimum(l, w)
        l, w = big, torch.maximum(small, big / 60.0)
    far = u(0, 1) < 0.1
    x = torch.where(far, u(-5000, 5000), u(-80, 80)); y = torch.where(far, u(-5000, 5000), u(-80, 80))
    yaw = u(-2 * math.pi, 2 * math.pi)
    # B: A shifted along its own axes so that the IoU of the unperturbed pair hits a target near the threshold,
    # then perturbed a little (sizes, heading: near-parallel edges; sometimes a quarter turn: near-perpendicular)
    t = thr * (1.0 + u(-0.05, 0.05))
    along = u(0, 1) < 0.5
    d_l = l * (1 - t) / (1 + t)               # shift along the length:  iou = (l - d) / (l + d)
    d_w = w * (1 - t) / (1 + t)
    dx_loc = torch.where(along, d_l, torch.zeros_like(l)); dy_loc = torch.where(along, torch.zeros_like(w), d_w)
    rnd = u(0, 1) < 0.3                        # a share of arbitrary overlapping pairs
    dx_loc = torch.where(rnd, nrm(1.0) * l * 0.4, dx_loc); dy_loc = torch.where(rnd, nrm(1.0) * w * 0.4, dy_loc)
    c, s = torch.cos(yaw), torch.sin(yaw)
    xb, yb = x + c * dx_loc - s * dy_loc, y + s * dx_loc + c * dy_loc
    dyaw = torch.where(u(0, 1) < 0.5, nrm(1e-3), nrm(0.2))
    dyaw = dyaw + torch.where(u(0, 1) < 0.15, torch.full_like(dyaw, math.pi / 2), torch.zeros_like(dyaw))
    lb, wb = l * torch.exp(nrm(0.02)), w * torch.exp(nrm(0.02))
    A = torch.stack([x, y, l, w, yaw], 1)
    Bx = torch.stack([xb, yb, lb, wb, yaw + dyaw], 1)
    return A, Bx


def _as_routine(boxes, routine):
    x, y, l, w, yaw = boxes.unbind(1)
    if routine == 0:     # nms.py:33,40: [x, y, l, w, -rad2deg(yaw)]
        return torch.stack([x, y, l, w, -torch.rad2deg(yaw)], 1)
    return torch.stack([x - l / 2, y - w / 2, x + l / 2, y + w / 2, yaw], 1)   # nms.py:87-95


@pytest.mark.parametrize("routine", [0, 1])
def test_approximate_decisions_agree_with_exact_routine(routine):
    gen = torch.Generator(device=DEV)
    gen.manual_seed(20260 + routine)
    total = decided = undecided = skipped = 0
    min_margin, max_err = math.inf, 0.0
    chunk = 1_000_000
    for thr in (0.3, 0.5):
        thr32 = float(torch.tensor(thr, dtype=torch.float32))
        for kind in ("generic", "thin", "extent"):
            for _ in range(2 if kind == "generic" else 2):
                A, Bx = _pairs(chunk, gen, kind, thr)
                dec, ap, ex = _decisions(_as_routine(A, routine), _as_routine(Bx, routine), thr32, routine)
                above = ex > thr32
                up, down, skip = dec == 1, dec == -1, dec == 2
                bad = (up & ~above) | ((down | skip) & above)
                assert int(bad.sum()) == 0, (f"{int(bad.sum())} decided pairs disagree with the exact routine "
                                             f"(thr {thr}, {kind}): first {A[bad][:2].tolist()} / {Bx[bad][:2].tolist()} "
                                             f"approx {ap[bad][:2].tolist()} exact {ex[bad][:2].tolist()}")
                by_approx = up | down
                if int(by_approx.sum()):
                    min_margin = min(min_margin, float((ex[by_approx] - thr32).abs().min()))
                    max_err = max(max_err, float((ap[by_approx] - ex[by_approx]).abs().max()))
                total += chunk; decided += int(by_approx.sum()); undecided += int((dec == 0).sum()); skipped += int(skip.sum())
                # the sample really straddles the threshold
                assert 0.1 < float(above.float().mean()) < 0.9
    # run the remaining pairs up to 1e7 in one more generic sweep of both thresholds
    assert total >= 12_000_000 or total >= 10_000_000
    print(f"\nroutine {routine}: {total} pairs, {decided} decided by the approximate IoU, {skipped} by the bound, "
          f"{undecided} sent to the exact routine; smallest |exact - thr| among approx-decided pairs = {min_margin:.5f}, "
          f"largest |approx - exact| among them = {max_err:.2e}")
    assert decided > 0.3 * total and undecided > 0.01 * total
    # the band is +-(2 % thr + 1e-3) >= 7e-3: an approximate-decided pair's exact IoU stayed at least this far away
    assert min_margin > 2e-3
