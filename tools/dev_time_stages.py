"""Dev helper: time rasterize and decode_compact alone (CUDA events, L2 flushed) at a bench shape."""
import sys, math
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
import bench
from tests import synth
from rv3d.math.range_view import pack_sweeps, rasterize_sweeps
from rv3d.nn.decoders.range_decoder import RangeDecoder
shape = sys.argv[1] if len(sys.argv) > 1 else "waymo"
B = 16
n, H, W, C, M, ident = bench.WORKLOADS[shape]
dev = torch.device("cuda:0")
sweeps, head, _ = bench.make_inputs(shape, B, 1000)
pts, las, cnt = [t.to(dev) for t in pack_sweeps(sweeps, dev)]
hd = {k: v.to(dev) for k, v in head.items()}
mapping = torch.arange(H, dtype=torch.int32, device=dev)
out = torch.empty((B, 7, H, W), dtype=torch.float32, device=dev)
ws = torch.empty(B * H * W * 8, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
dec = RangeDecoder(True, True, *bench.SBR)
pp = dict(bench.PP, nms_mode="HARD"); tasks = {0: [f"c{i}" for i in range(C)]}
ms = {1: {"cart": hd["cart"], "mask": hd["mask"], 0: {"logits": hd["logits"], "regressands": hd["regressands"]}}}
def timeit(fn, reps=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))
t_r = timeit(lambda: rasterize_sweeps(pts, las, cnt, mapping, synth.LIDAR_OFFSET, H, W, out=out, workspace=ws))
cand = dec.candidates(ms, pp, tasks); ncand = cand.count()
t_d = timeit(lambda: dec.candidates(ms, pp, tasks))
rb, db = bench.algorithmic_bytes(shape, B, ncand)
print(f"{shape} B={B}: rasterize {t_r*1e3:.1f} us ({rb/t_r/1e6:.0f} GB/s)  decode_compact {t_d*1e3:.1f} us ({db/t_d/1e6:.0f} GB/s)  "
      f"combined {(rb+db)/(t_r+t_d)/1e6:.0f} GB/s = {(rb+db)/(t_r+t_d)/1e6/6451.2:.3f} of measured HBM peak; survivors {ncand}")
