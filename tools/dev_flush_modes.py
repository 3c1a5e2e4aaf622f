"""Dev helper: the roofline pair (rasterize + decode_compact, forked and serial graphs) under different L2 flush
protocols.  A 256 MiB memset leaves ~L2-capacity of DIRTY lines behind; their write-back then competes with the timed
kernels for DRAM bandwidth.  Modes: zero (memset only), zero+read (memset, then a read pass over a second buffer so the
L2 holds clean lines when the timed region starts), read (read pass only), none (back-to-back replays)."""
import sys
from pathlib import Path
import numpy as np, torch
ROOT = Path(__file__).resolve().parents[1]
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]
import bench

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
B = 16
hp = bench.HotPath("waymo", B, dev, "HARD")
box = {}
def stage_decode():
    box["c"] = hp.dec.candidates(hp.ms_of(hp.hd), hp.pp, hp.tasks)
def serial():
    hp.rasterize(); stage_decode()
def forked():
    cur = torch.cuda.current_stream(dev)
    hp.side.wait_stream(cur)
    with torch.cuda.stream(hp.side):
        hp.rasterize()
    stage_decode()
    cur.wait_stream(hp.side)
serial(); torch.cuda.synchronize()
ncand = box["c"].count()
graphs = {"raster": bench.capture(hp.rasterize, dev)[0], "decode": bench.capture(stage_decode, dev)[0],
          "serial": bench.capture(serial, dev)[0], "forked": bench.capture(forked, dev)[0]}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
clean = torch.zeros(256 << 17, dtype=torch.int64, device=dev)   # 256 MiB, read-only
rb, db = bench.algorithmic_bytes("waymo", B, ncand)

def pre(mode):
    if mode in ("zero", "zero+read"):
        flush.zero_()
    if mode in ("read", "zero+read"):
        clean.sum()

for mode in ("zero", "zero+read", "read", "none"):
    for name, g in graphs.items():
        for _ in range(3):
            g.replay()
        ts = []
        for _ in range(20):
            pre(mode)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts = np.array(ts)
        nbytes = {"raster": rb, "decode": db}.get(name, rb + db)
        print(f"{mode:10s} {name:7s} mean {ts.mean()*1e3:7.1f} us  median {np.median(ts)*1e3:7.1f}  min {ts.min()*1e3:7.1f}  "
              f"-> {nbytes / ts.mean() / 1e6:6.0f} GB/s = {nbytes / ts.mean() / 1e6 / 6451.2:.3f}", flush=True)
