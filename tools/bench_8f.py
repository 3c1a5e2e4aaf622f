"""Device timing of the SURVEY 8f "next" rows' kernels at production sizes, against the measured HBM peak, with the
oracle's CPU time on a bounded sample beside it.  One JSON line per kernel; run on a B200:

    python tools/bench_8f.py > gpurun_out/bench_8f.jsonl

Inputs are larger than L2 or the L2 is flushed between iterations (stated per line)."""
from __future__ import annotations

import json
import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "range-view-3d-detection_b200"))
from tests import synth  # noqa: E402

DEV = torch.device("cuda:0")
PEAK = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text()).get("hbm_gbs", 6650.0)) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
FLUSH = None


def timed(fn, iters=20, warmup=3):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)
    for _ in range(warmup):
        fn()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        FLUSH.zero_()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.mean([a.elapsed_time(b) for a, b in ev]))


def report(name, ms, alg_bytes, units, unit_name, cpu_s=None, cpu_units=None, note=""):
    gbs = alg_bytes / (ms * 1e-3) / 1e9
    line = {"kernel": name, "ms": round(ms, 4), "algorithmic_bytes": int(alg_bytes), "achieved_gbs": round(gbs, 1),
            "hbm_peak_gbs": PEAK, "frac": round(gbs / PEAK, 3), f"{unit_name}_per_s": units / (ms * 1e-3), "l2": "flushed between iterations",
            "note": note}
    if cpu_s is not None:
        line["cpu_oracle"] = {f"{unit_name}_per_s": cpu_units / cpu_s, "sample": f"{cpu_units} {unit_name}, 1 process (numpy / torch CPU)"}
        line["speedup_vs_cpu_oracle"] = (units / (ms * 1e-3)) / (cpu_units / cpu_s)
    print(json.dumps(line), flush=True)


def main():
    from oracle import assign_oracle, av2_prep
    import oracle
    from rv3d.converters.av2 import utils as cu
    from rv3d.math.ops.assignment import box_iou_rotated, compute_classification_targets
    from rv3d.math.ops.coding import decode_range_view
    from rv3d.prototype.loader import range_view_inputs, rasterize_inputs
    from rv3d.math.range_view import pack_sweeps, rasterize_sweeps

    B, H, W = 16, 64, 2650
    # ---- row 1: loader inputs --------------------------------------------------------------
    image = torch.randn((B, 7, H, W), device=DEV)
    image[:, 2] = image[:, 2].abs() * (torch.rand((B, H, W), device=DEV) < 0.7)
    for stride in (1, 4):
        ms = timed(lambda: range_view_inputs(image, dataset_name="waymo", x_stride=stride, mode="constant"))
        wo = range_view_inputs(image, dataset_name="waymo", x_stride=stride, mode="constant")[0].shape[-1]
        alg = B * H * (5 * 4 * W / stride + (8 * 4 + 1) * wo)       # 5 distinct planes read at the kept columns, 8 planes + mask written
        cpu_img = image[:1].cpu()
        t0 = time.perf_counter()
        f, c, m = cpu_img[0, [6, 2, 3, 4, 5]].clone(), cpu_img[0, 3:6].clone(), cpu_img[0, 2:3] > 0
        f[0] = f[0].tanh()
        f *= m
        pad = 0
        _ = (f[..., ::stride].contiguous(), c[..., ::stride].contiguous(), m[..., ::stride].contiguous())
        cpu = time.perf_counter() - t0
        report(f"range_view_inputs(stride={stride})", ms, alg, B, "sweeps", cpu, 1, "row 1: features/cart/mask assembly + subsample_range_view")

    # raw sweeps -> network inputs: rasterize_sweeps + range_view_inputs (two passes over the 76 MB image) vs the fused entry
    sweeps = [synth.make_points(180_000, H, 1000 + s_) for s_ in range(B)]
    pts, las_, cnt = [t.to(DEV) for t in pack_sweeps(sweeps, DEV)]
    mapping = torch.arange(H, dtype=torch.int32, device=DEV)
    ws = torch.empty(B * H * W * 8, dtype=torch.uint8, device=DEV)
    img_out = torch.empty((B, 7, H, W), dtype=torch.float32, device=DEV)
    for stride in (1, 4):
        alg = B * (180_000 * 17 + H * ((W + 6) // stride) * (8 * 4 + 1))     # points + laser bytes in, 5 + 3 planes + mask out
        ms2 = timed(lambda: range_view_inputs(rasterize_sweeps(pts, las_, cnt, mapping, synth.LIDAR_OFFSET, H, W, out=img_out, workspace=ws),
                                              dataset_name="waymo", x_stride=stride, mode="circular"))
        report(f"rasterize_sweeps + range_view_inputs(stride={stride})", ms2, alg, B, "sweeps", note="row 1: the unfused pair, bytes of the fused form")
        ms1 = timed(lambda: rasterize_inputs(pts, las_, cnt, mapping, synth.LIDAR_OFFSET, H, W, dataset_name="waymo", x_stride=stride,
                                             mode="circular", workspace=ws))
        report(f"rasterize_inputs(stride={stride})", ms1, alg, B, "sweeps", note="row 1: rasterizer with the loader assembly fused into its resolve pass")

    # ---- row 2: sweep preparation -----------------------------------------------------------
    n = 16 * 180_000
    ts, quat, trans = synth.make_pose_table(3000, seed=9)
    xyz, off, *_ = synth.make_raw_sweep(n, seed=10)
    dx, do = torch.from_numpy(xyz).to(DEV), torch.from_numpy(off).to(DEV)
    t0n = int(ts[1500])
    s = 180_000
    t0 = time.perf_counter(); av2_prep.unmotion_compensate(xyz[:s], off[:s], t0n, ts, quat, trans); cpu = time.perf_counter() - t0
    ms = timed(lambda: cu.unmotion_compensate(dx, do, t0n, ts, quat, trans))
    report("unmotion_compensate (per-call table)", ms, n * (32 + 25), n, "points", cpu, s,
           "row 2: the reference's signature: pose table uploaded and its Slerp intervals prepared on every call, one host read (dropped-row count)")
    table = cu.PoseTable(ts, quat, trans, device=DEV)
    ms = timed(lambda: cu.unmotion_compensate(dx, do, t0n, table))
    report("unmotion_compensate (resident PoseTable)", ms, n * (32 + 25), n, "points", cpu, s,
           "row 2: one PoseTable per log; fp64-issue-bound (~330 fp64 instructions per point), one host read (dropped-row count)")
    rot = av2_prep.quat_to_matrix(np.array([0.0012, -0.0031, 0.0052, 0.99998]))
    tr = np.array([1.35, 0.0, 1.64])
    ms = timed(lambda: cu.sensor_from_egovehicle(dx, rot, tr))
    t0 = time.perf_counter(); av2_prep.sensor_from_egovehicle(xyz[:s], rot, tr); cpu = time.perf_counter() - t0
    report("sensor_from_egovehicle", ms, n * 48, n, "points", cpu, s, "row 2")
    las = torch.randint(0, 64, (n,), device=DEV)
    ms = timed(lambda: cu.correct_laser_numbers(las, "a", 64, log_ids=("a",)))
    ln = las[:s].cpu().numpy()
    t0 = time.perf_counter(); av2_prep.correct_laser_numbers(ln, True, 64); cpu = time.perf_counter() - t0
    report("correct_laser_numbers", ms, n * 16, n, "points", cpu, s, "row 2: includes the out-of-range flag read (host sync, numpy's IndexError)")
    ms = timed(lambda: cu.correct_laser_numbers(las, "a", 64, log_ids=("a",), validate=False))
    report("correct_laser_numbers(validate=False)", ms, n * 16, n, "points", cpu, s, "row 2: no host read (out-of-table rows come back as -1)")

    # ---- row 4: training-time callers --------------------------------------------------------
    head = synth.make_head_outputs(B, 3, H, W, seed=1, n_objects=32)
    reg, cart = head["regressands"].to(DEV), head["cart"].to(DEV)
    ms = timed(lambda: decode_range_view(reg, cart, True))
    t0 = time.perf_counter(); oracle.decode_range_view(head["regressands"][:1], head["cart"][:1], True); cpu = time.perf_counter() - t0
    report("decode_range_view (dense)", ms, B * H * W * (11 * 4 + 7 * 4), B, "sweeps", cpu, 1, "row 4: called twice per task by compute_classification_targets")
    cub = synth.make_nms_candidates(1, 1_000_000, 1, 5000, seed=3)[0][0]
    a5 = cub[:, [0, 1, 3, 4, 6]].contiguous().to(DEV)
    b5 = (a5 + 0.3 * torch.randn_like(a5)).contiguous()
    ms = timed(lambda: box_iou_rotated(a5, b5, aligned=True))
    t0 = time.perf_counter(); assign_oracle.box_iou_rotated(a5[:100_000].cpu(), b5[:100_000].cpu(), aligned=True); cpu = time.perf_counter() - t0
    report("box_iou_rotated(aligned)", ms, a5.shape[0] * 44, a5.shape[0], "pairs", cpu, 100_000, "row 4: compute-bound (bit-exact rotated IoU), bytes for reference only")
    d = synth.make_assignment_inputs(4, 3, H, W, seed=5, n_instances=150)
    dv = {k: v.to(DEV) for k, v in d.items()}
    cfg = dict(affinity_fn="bev", enable_azimuth_invariant_targets=True, k=8, normalize_affinities=False, sigma=1.0)
    d1 = {k: v[:1] for k, v in d.items()}
    from rv3d.math.ops import assignment as A
    io_bytes = 4 * H * W * ((8 + 8 + 3) * 4 + 8 + 8 + 1 + (3 + 1) * 4 + 2)      # every input once + every output once
    for tag, c in (("BEV k=8", cfg), ("GAUSSIAN k=inf (production config)", dict(cfg, affinity_fn="gaussian", k=float("inf"), sigma=0.75))):
        t0 = time.perf_counter(); assign_oracle.compute_classification_targets(d1["input"], d1["target"], d1["labels"], d1["cart"], c, d1["mask"], d1["panoptics"], 3); cpu = time.perf_counter() - t0
        args = (dv["input"], dv["target"], dv["labels"], dv["cart"], c, dv["mask"], dv["panoptics"], 3)
        ms = timed(lambda: A._compute_classification_targets_composed(*args[:4], dict(c), *args[5:], str(c["affinity_fn"]).upper(), A._k_slots(c["k"])), iters=10)
        report(f"compute_classification_targets composed, {tag}", ms, io_bytes, 4, "sweeps", cpu, 1,
               "row 4: round-1 form (2 dense decodes, torch gathers, IoU, one stable sort, scatters)")
        ms = timed(lambda: compute_classification_targets(*args), iters=10)
        report(f"compute_classification_targets fused, {tag}", ms, io_bytes, 4, "sweeps", cpu, 1,
               "row 4: ONE call (foreground-only decode, atomic top-k slots) + one host read of the largest instance id")
        ms = timed(lambda: compute_classification_targets(*args, max_instances=256), iters=10)
        report(f"compute_classification_targets fused, max_instances given, {tag}", ms, io_bytes, 4, "sweeps", cpu, 1,
               "row 4: ONE call, no host read")


if __name__ == "__main__":
    main()
