"""Dev helper: one bench step with the NMS kernel's per-segment debug print (build with
RV3D_NVCC_DEFS=-DRV3D_NMS_DEBUG_PRINT)."""
import sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
import bench
dev = torch.device("cuda:0")
hp = bench.HotPath("waymo", 16, dev, sys.argv[1] if len(sys.argv) > 1 else "HARD")
hp.decode(stats=None); torch.cuda.synchronize()
print("---- step with stats", flush=True)
det = hp.decode(stats=hp.stats); torch.cuda.synchronize()
print("detections", det.wait())
