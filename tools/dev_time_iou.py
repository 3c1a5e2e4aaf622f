"""Dev helper: throughput of the exact rotated IoU device routine (through rv3d_iou3d_aligned)."""
import sys, math
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
from rv3d.math.ops.iou import iou_3d_axis_aligned
rng = np.random.default_rng(0)
n = 2_000_000
a = np.empty((n, 7), np.float32)
a[:, :2] = rng.uniform(-50, 50, (n, 2)); a[:, 2] = 0
a[:, 3:6] = np.exp(rng.normal(0.8, 0.3, (n, 3))); a[:, 6] = rng.uniform(-math.pi, math.pi, n)
b = a.copy(); b[:, :2] += rng.normal(0, 1.0, (n, 2)).astype(np.float32); b[:, 6] += rng.normal(0, 0.3, n).astype(np.float32)
A, B = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
for _ in range(3): iou_3d_axis_aligned(A, B)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): i3, bev = iou_3d_axis_aligned(A, B)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{n} aligned rotated IoUs: {ms:.3f} ms -> {n/ms/1e6:.2f} G IoU/s, {ms*1e6/n:.2f} ns/IoU ; frac>0.3 {(bev>0.3).float().mean().item():.2f}")
