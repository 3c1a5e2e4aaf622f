"""Dev helper: time the rasterizer alone at Waymo shape (B=16) with CUDA events."""
import sys
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
from tests import synth

dev = torch.device("cuda:0")
B, N, H, W = 16, 180_000, 64, 2650
sweeps = [synth.make_points(N, H, s) for s in range(B)]
pts, las, cnt = [t.to(dev) for t in pack_sweeps(sweeps, dev)]
mapping = torch.arange(H, dtype=torch.int32, device=dev)
out = torch.empty((B, 7, H, W), dtype=torch.float32, device=dev)
ws = torch.empty(B * H * W * 8, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(5):
    rasterize_sweeps(pts, las, cnt, mapping, synth.LIDAR_OFFSET, H, W, out=out, workspace=ws)
ts = []
for _ in range(20):
    flush.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rasterize_sweeps(pts, las, cnt, mapping, synth.LIDAR_OFFSET, H, W, out=out, workspace=ws)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
bytes_alg = B * (N * 17 + 7 * H * W * 4)
print(f"raster B={B}: median {np.median(ts)*1e3:.1f} us  min {ts.min()*1e3:.1f} us  "
      f"alg {bytes_alg/1e6:.1f} MB -> {bytes_alg/np.median(ts)/1e6:.0f} GB/s")
