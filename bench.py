#!/usr/bin/env python
"""bench.py -- sweeps/sec of the rasterize -> decode -> NMS path on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--nms-mode HARD|WEIGHTED]
                    [--no-extras] [--global-batch G]

One "step" = one pass of the hot path over one batch of synthetic sweeps: rasterize B raw sweeps
(N points each -> 64 x W range image), then RangeDecoder.decode of B sets of dense head outputs (sigmoid / max /
threshold / sample_by_range / box decode / NMS).  Workload at N=1: BASELINE.json configs[1], Waymo shape, B=16.
With N>1 (torchrun, one rank per GPU) every rank runs its own B sweeps (weak scaling) and the step ends with the
path's one exchange, a gather of the detections, fused into the NMS pack kernel.

How the numbers are taken:
  value      the step's public calls (rv3d.math.range_view.rasterize_sweeps + RangeDecoder.decode_async) are captured
             ONCE in a CUDA graph -- the step has no host read, so it replays as is -- and replayed K times with
             inputs resident in HBM: steady-state throughput with --pipeline-depth step graphs in flight on as many
             streams (CUDA events from the first launch to the last completion); `single_stream` = one step at a time,
             per-step CUDA events, 256 MiB L2 flush between steps (the step's latency).
  roofline   rasterize + decode_compact replayed as ONE graph the way a step runs the pair (rasterizer on a forked
             stream), CUDA events around the replay, L2 flushed before each; achieved = SURVEY 8d algorithmic bytes /
             that time, peak = MEASURED_PEAKS.json; the serial form, each stage alone and the same sweeps in a
             sensor's firing order are reported beside it.
  e2e        the same calls, eager, from pinned HOST buffers: H2D of every input (double-buffered on a copy
             stream), the step, D2H of the detections (on a second copy stream), every step; wall clock.
  extras     BASELINE configs 1, 3, 4 (and 5 at N>1) and the reference's own batch-1 latency protocol
             (tools/benchmark.py:231-238), a few steps each, in the same JSON line under "extra".
`--impl reference` times the reference's CPU path (oracle port: /root/reference does not exist on the GPU box, and
its third-party natives are not installable) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "range-view-3d-detection_b200")]

from tests import synth  # noqa: E402

METRIC = "sweeps/sec (rasterize+decode+NMS)"
WORKLOADS = {
    # name: (points per sweep, H, W, classes, objects per sweep, identity row map?)
    "waymo": (180_000, 64, 2650, 3, 96, True),
    "av2": (100_000, 64, 1800, 26, 64, False),
    # BASELINE config 4: the same sweeps rasterized / decoded at three widths
    "w900": (100_000, 64, 900, 3, 64, True),
    "w1800": (100_000, 64, 1800, 3, 64, True),
    "w3600": (100_000, 64, 3600, 3, 64, True),
}
PP = {"num_pre_nms": 50000, "num_post_nms": 1000, "nms_threshold": 0.3, "min_confidence": 0.1}  # range_view.yaml:43-47
SBR = ([0, 15, 30], [15, 30, math.inf], [8, 2, 1])                                              # range_view.yaml:133-135
FP_RATE = 0.95   # fraction of valid background pixels that fire (S ~ 20 % of K pass 0.1)


def make_inputs(shape: str, batch: int, seed0: int, fp_rate: float = None):
    n, H, W, C, M, ident = WORKLOADS[shape]
    sweeps = [synth.make_points(n, H, seed0 + s) for s in range(batch)]
    head = synth.make_head_outputs(batch, C, H, W, seed=seed0, n_objects=M, fp_rate=FP_RATE if fp_rate is None else fp_rate, distinct_scores=False)
    mapping = np.arange(H) if ident else None
    return sweeps, head, mapping


def firing_order(sweep, n_azimuth_bins: int):
    """The same points in the order a spinning lidar emits them: azimuth step by azimuth step, all lasers per step
    (synth.make_points draws laser and azimuth independently per point, i.e. a shuffled sweep: the worst case for the
    rasterizer's scattered atomics and gathers)."""
    xyz, inten, laser = sweep
    rel = xyz.astype(np.float64) - synth.LIDAR_OFFSET
    col = np.floor((np.arctan2(rel[:, 1], rel[:, 0]) + math.pi) / (2 * math.pi) * n_azimuth_bins).astype(np.int64)
    order = np.lexsort((laser, col))
    return xyz[order], inten[order], laser[order]


def algorithmic_bytes(shape: str, batch: int, survivors: int, head_bytes: int = 4):
    """rasterize reads N*(16+1) B and writes 7*H*W*4 B per sweep; decode reads
    H*W*(4*(C+8+3)+1) B per sweep (dense count: every input once) and writes 40 B per survivor.
    (`head_bytes` = 2 with f16 heads: logits and regressands in half precision, cart stays float32.)"""
    n, H, W, C, _, _ = WORKLOADS[shape]
    raster = batch * (n * 17 + 7 * H * W * 4)
    decode = batch * (H * W * (head_bytes * (C + 8) + 4 * 3 + 1)) + 40 * survivors
    return raster, decode


def ncu_traffic(shape: str, batch: int):
    """dram__bytes_read.sum + dram__bytes_write.sum of scatter + resolve + decode_compact per step, from the committed
    `ncu --set full` capture of the default workload (profiles/traffic.json, written by tools/ncu_summary.py from the
    .ncu-rep); other workloads were not captured -> None."""
    try:
        t = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        return int(t["bytes_per_step"]) if (t.get("shape"), t.get("batch")) == (shape, batch) else None
    except (OSError, ValueError, KeyError):
        return None


# --------------------------------------------------------------------------------------- #
# clocks                                                                                   #
# --------------------------------------------------------------------------------------- #
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, enabled: bool = True):
        self.index, self.proc, self.lines, self.enabled = index, None, [], enabled

    def __enter__(self):
        if not self.enabled:   # one sampler per job (rank 0): nvidia-smi polls take a driver-wide lock
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6 or not f[0].isdigit():
                continue
            sm.append(int(f[0])); mx = max(mx, int(f[1]))
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------------------- #
# the CPU arm (oracle port of the reference's CPU path)                                    #
# --------------------------------------------------------------------------------------- #
def cpu_step(sweep, head1, mapping, shape: str, nms_mode: str, pool: ThreadPoolExecutor):
    """One sweep through the reference's CPU path, restated (oracle/): numpy rasterizer + serial z-buffer,
    torch-CPU decode on all host threads, C rotated NMS with one host thread per class."""
    import oracle
    _, H, W, C, _, _ = WORKLOADS[shape]
    t0 = time.perf_counter()
    oracle.build_range_view(sweep[0], sweep[1], sweep[2], mapping if mapping is not None else oracle.ROW_MAPPING_64,
                            synth.LIDAR_OFFSET, num_lasers=H, width=W, n_azimuth_bins=W)
    t1 = time.perf_counter()
    ms = {1: {"cart": head1["cart"], "mask": head1["mask"], 0: {"logits": head1["logits"], "regressands": head1["regressands"]}}}
    tasks = {0: [f"c{i}" for i in range(C)]}
    params, scores, cats = oracle.range_decoder_decode(ms, PP, tasks, True, True, *SBR, return_candidates=True)
    t2 = time.perf_counter()
    live = scores[0] >= PP["min_confidence"]
    cu, sc, ca = params[0, live], scores[0, live], cats[0, live]
    fn = oracle.hard_multiclass_nms if nms_mode == "HARD" else oracle.weighted_multiclass_nms

    def one_class(j):
        m = ca == j
        return fn(cu[m], sc[m], ca[m], PP["nms_threshold"], PP["num_pre_nms"], PP["num_post_nms"])[1].shape[0]

    kept = sum(pool.map(one_class, torch.unique(ca).tolist()))
    t3 = time.perf_counter()
    return {"rasterize_ms": (t1 - t0) * 1e3, "decode_ms": (t2 - t1) * 1e3, "nms_ms": (t3 - t2) * 1e3,
            "total_s": t3 - t0, "kept": kept, "survivors": int(live.sum())}


def cpu_baseline_record(stages, cores: int, shape: str):
    """`stages`: list of cpu_step results -> the cpu_baseline object (value = sweeps/s of the whole path; the NMS share is
    stated because the restated detectron2 CPU NMS dominates it -- a path the reference itself runs on the GPU)."""
    tot = float(np.mean([s["total_s"] for s in stages]))
    sm = {k: float(np.mean([s[k] for s in stages])) for k in ("rasterize_ms", "decode_ms", "nms_ms")}
    return {"value": 1.0 / tot, "unit": "sweeps/s", "cores": cores, "kind": "port",
            "sample": f"1 {shape}-shaped sweep per step ({WORKLOADS[shape][0]} pts, 64x{WORKLOADS[shape][2]}, "
                      f"{stages[-1]['survivors']} candidates >= 0.1): numpy rasterize + serial z-buffer (1 thread), torch-CPU "
                      f"decode ({cores} threads), C rotated NMS (1 thread per class, greedy scan stopped at num_post_nms kept)",
            "stage_ms": sm, "nms_share": sm["nms_ms"] / (tot * 1e3),
            "sweeps_per_s_rasterize_decode_only": 1e3 / (sm["rasterize_ms"] + sm["decode_ms"])}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle  # noqa: F401  (builds the C part)
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sweeps, head, mapping = make_inputs(args.shape, 1, 1000, args.fp_rate)
    pool = ThreadPoolExecutor(max_workers=cores)
    for _ in range(args.warmup):
        cpu_step(sweeps[0], head, mapping, args.shape, args.nms_mode, pool)
    t0 = time.perf_counter()
    stages = [cpu_step(sweeps[0], head, mapping, args.shape, args.nms_mode, pool) for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    value = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            # the CUDA arm's config; each step here is a bounded sample of it (1 of the batch's sweeps, see `sample`)
            "config": dict(workload_config(args, args.batch), l2="n/a (host arm)", sample_per_step="1 sweep"),
            "cpu_baseline": dict(cpu_baseline_record(stages, cores, args.shape), value=value),
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, batch):
    n, H, W, C, M, _ = WORKLOADS[args.shape]
    return {"workload": f"{args.shape}-shaped synthetic sweeps: rasterize ({n} pts -> {H}x{W}x7) + RangeDecoder.decode "
                        f"({C} classes, sample_by_range [8,2,1], azimuth-invariant) + {args.nms_mode} rotated NMS "
                        f"(pre {PP['num_pre_nms']}, post {PP['num_post_nms']}, iou {PP['nms_threshold']}, conf {PP['min_confidence']})",
            "batch_per_gpu": batch, "points_per_sweep": n, "height": H, "width": W, "classes": C,
            "objects_per_sweep": M, "fp_rate": FP_RATE if getattr(args, "fp_rate", None) is None else args.fp_rate, "nms_mode": args.nms_mode,
            "l2": "256 MiB L2 flush between timed steps (outside the per-step CUDA events)"}


def bind_to_gpu_numa_node(index: int) -> str:
    """Pin this process to the CPUs NVML reports as local to GPU `index` (torchrun does not bind ranks): the pinned host
    buffers of the e2e leg are then first-touched on the GPU's own NUMA node.  Best effort; returns what was done."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus and len(cpus) < len(os.sched_getaffinity(0)):
            os.sched_setaffinity(0, cpus)
            return f"bound to {len(cpus)} CPUs local to GPU {index}"
        return "GPU-local CPU set == current affinity"
    except Exception as exc:   # noqa: BLE001
        return f"not bound ({type(exc).__name__})"


# --------------------------------------------------------------------------------------- #
# the CUDA arm                                                                             #
# --------------------------------------------------------------------------------------- #
class HotPath:
    """One configured instance of the path: resident inputs + the public calls of a step (nothing else)."""

    def __init__(self, shape, batch, dev, nms_mode, head_dtype="f32", fp_rate=None, gen_batch=None, peer=None, rank=0,
                 inputs=None):
        from rv3d.constants import ROW_MAPPING_64
        from rv3d.math.range_view import pack_sweeps
        from rv3d.nn.decoders.range_decoder import RangeDecoder
        self.shape, self.B, self.dev, self.peer, self.rank = shape, batch, dev, peer, rank
        n, H, W, C, M, ident = WORKLOADS[shape]
        self.H, self.W, self.C = H, W, C
        g = min(gen_batch or batch, batch)
        # weak scaling: every rank gets the SAME synthetic sweeps, so the per-GPU work is identical at every N (NMS time is
        # data dependent; with different seeds the max over ranks would measure the unluckiest seed, not the scaling)
        self.sweeps, head, self.mapping = inputs if inputs is not None else make_inputs(shape, g, 1000, fp_rate)
        if head_dtype == "f16":   # the reference's own eval_precision = 16 operating point (range_view.yaml:26, detector.py:329-333)
            head = dict(head, logits=head["logits"].half(), regressands=head["regressands"].half())
        self.head_host = head
        self.pts_h, self.las_h, self.cnt_h = pack_sweeps(self.sweeps, dev, pin=True)
        self.head_h = {k: v.pin_memory() for k, v in head.items()}
        rep = batch // g

        def tile(t):
            t = t.to(dev)
            return t.repeat(rep, *([1] * (t.dim() - 1))) if rep > 1 else t

        self.pts, self.las, self.cnt = tile(self.pts_h), tile(self.las_h), tile(self.cnt_h)
        self.hd = {k: tile(v) for k, v in self.head_h.items()}
        self.row_map = torch.as_tensor((np.arange(H) if ident else ROW_MAPPING_64).astype(np.int32), device=dev)
        self.pp = dict(PP, nms_mode=nms_mode)
        self.tasks = {0: [f"c{i}" for i in range(C)]}
        self.dec = RangeDecoder(True, True, *SBR)
        self.image = torch.empty((batch, 7, H, W), dtype=torch.float32, device=dev)
        self.rws = torch.empty(batch * H * W * 8, dtype=torch.uint8, device=dev)
        self.stats = torch.zeros(24, dtype=torch.int64, device=dev)
        self.side = torch.cuda.Stream(dev, priority=0)

    @staticmethod
    def ms_of(h):
        return {1: {"cart": h["cart"], "mask": h["mask"], 0: {"logits": h["logits"], "regressands": h["regressands"]}}}

    # ---- the calls a user makes ----
    def rasterize(self, p=None, l=None, c=None):
        from rv3d.math.range_view import rasterize_sweeps
        return rasterize_sweeps(self.pts if p is None else p, self.las if l is None else l, self.cnt if c is None else c,
                                self.row_map, synth.LIDAR_OFFSET, self.H, self.W, out=self.image, workspace=self.rws)

    def decode(self, h=None, stats=None):
        gather = self.peer.flagged(self.rank * self.B) if self.peer is not None else None
        return self.dec.decode_async(self.ms_of(self.hd if h is None else h), self.pp, self.tasks, gather=gather, stats=stats)

    def step(self, p=None, l=None, c=None, h=None, stats=None):
        # The two halves of a step are independent (the backbone sits between them in the real model: the rasterizer
        # prepares the NEXT batch's input while the decoder works on this batch's head outputs), so the rasterizer goes to
        # a forked stream behind the dense decode kernel (RangeDecoder.dense_done) and the step joins it at the end: it runs
        # on the ~100 SMs the 48-CTA suppression kernel leaves idle.  Under capture this is a fork / join inside the graph.
        cur = torch.cuda.current_stream(self.dev)
        self.side.wait_stream(cur)
        det = self.decode(h, stats)
        self.side.wait_event(self.dec.dense_done)      # the whole-GPU decode kernel first, then the rasterizer fills the
        with torch.cuda.stream(self.side):             # SMs the 48-CTA suppression kernel leaves idle
            self.rasterize(p, l, c)
        cur.wait_stream(self.side)
        return det

    def step_serial(self, p=None, l=None, c=None, h=None, stats=None):
        self.rasterize(p, l, c)
        return self.decode(h, stats)


def capture(fn, dev):
    """fn() enqueues work through the public API; -> (graph, fn's return value).  The workspaces were sized by an eager
    warm-up call before: nothing allocates from the default pool while capturing."""
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream(dev, priority=-1)   # the forked rasterizer stream has the lower priority: its blocks fill the SMs the main branch leaves idle
    s.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.graph(g, stream=s):
        out = fn()
    torch.cuda.current_stream(dev).wait_stream(s)
    return g, out


def replay_timed(graphs, steps, flush, barrier=None):
    """Replays the graph(s) back to back `steps` times; events around each one on the current stream, L2 flush between
    steps.  -> (per-step ms array per graph, wall seconds per step)."""
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(len(graphs) + 1)] for _ in range(steps)]
    if barrier:
        barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        if flush is not None:
            flush.zero_()
        ev[k][0].record()
        for i, g in enumerate(graphs):
            g.replay()
            ev[k][i + 1].record()
    if barrier:
        barrier()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) / steps
    return [np.array([e[i].elapsed_time(e[i + 1]) for e in ev]) for i in range(len(graphs))], wall


def run_extras(args, dev, flush, cores_all):
    """BASELINE configs 1, 3, 4 and the batch-1 latency protocol: small, bounded legs, each through the public API."""
    from rv3d.math.ops.nms import batched_multiclass_nms
    steps = max(3, min(args.steps, 10))
    extra = {}

    def graph_leg(hp):
        hp.step(); hp.step()
        torch.cuda.synchronize()
        g, det = capture(hp.step, dev)
        for _ in range(3):
            g.replay()
        (t,), _ = replay_timed([g], steps, flush)
        return float(np.mean(t)), int(det.wait())

    # ---- config 1: AV2-shaped sweeps, standard rotated NMS; CPU figure beside it ----
    hp = HotPath("av2", 16, dev, "HARD", fp_rate=args.fp_rate)
    ms, ndet = graph_leg(hp)
    rec = {"workload": "av2-shaped (100000 pts, 64x1800, 26 classes) batch 16, HARD NMS", "ms_per_step": ms,
           "sweeps_per_s": 16 / (ms * 1e-3), "detections_per_step": ndet, "steps": steps}
    if not args.no_cpu_baseline:
        import oracle  # noqa: F401
        os.sched_setaffinity(0, cores_all)
        torch.set_num_threads(len(cores_all))
        pool = ThreadPoolExecutor(max_workers=len(cores_all))
        h1 = {k: v[:1] for k, v in hp.head_host.items()}
        cpu_step(hp.sweeps[0], h1, hp.mapping, "av2", "HARD", pool)
        r = cpu_step(hp.sweeps[0], h1, hp.mapping, "av2", "HARD", pool)
        rec["cpu_reference_port"] = dict(cpu_baseline_record([r], len(cores_all), "av2"))
    extra["config1_av2_hard"] = rec
    del hp

    # ---- config 3: weighted NMS stress, 200 k candidates of one class in one sweep ----
    cub, sc, ca = synth.make_nms_candidates(1, 200_000, 1, 256, seed=123, spread=75.0, frac_clustered=0.97)
    sc = 0.1 + 0.9 * sc
    cub, sc, ca = cub.to(dev), sc.to(dev), ca.to(dev)
    for mode in ("WEIGHTED", "HARD"):
        for _ in range(3):
            out = batched_multiclass_nms(cub, sc, ca, 200_000, 1000, 0.3, 0.1, mode)
        ts = []
        for _ in range(steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = batched_multiclass_nms(cub, sc, ca, 200_000, 1000, 0.3, 0.1, mode)   # includes its one host wait
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        extra[f"config3_stress_200k_{mode.lower()}"] = {
            "workload": "1 sweep, 1 class, 200000 candidates >= 0.1 around 256 objects, num_pre_nms 200000, num_post_nms 1000",
            "ms_per_call": float(np.median(ts)) * 1e3, "candidates_per_s": 200_000 / float(np.median(ts)), "kept": int(out[1].shape[0]),
            "protocol": "sync-bracketed wall clock around batched_multiclass_nms (public call)",
            "torchex": "TorchEx is not installable here (no network): keep-set / merged rows are compared against the "
                       "oracle in tests/test_gpu_nms.py::test_config3_weighted_stress_200k; the reference kernel would "
                       "build two 200000 x 3125 u64 masks (10 GB) and scan them on the host"}
    del cub, sc, ca

    # ---- config 4: multi-resolution range images 64 x {900, 1800, 3600}, batch 64 ----
    base = None
    for name in ("w900", "w1800", "w3600"):
        if base is None:
            base = [synth.make_points(WORKLOADS[name][0], 64, 1000 + s) for s in range(8)]
        n, H, W, C, M, _ = WORKLOADS[name]
        head = synth.make_head_outputs(8, C, H, W, seed=1000, n_objects=M, fp_rate=FP_RATE if args.fp_rate is None else args.fp_rate,
                                       distinct_scores=False)
        hp = HotPath(name, 64, dev, "HARD", gen_batch=8, inputs=(base, head, np.arange(H)))
        ms, ndet = graph_leg(hp)
        extra[f"config4_{name}"] = {"workload": f"{n} pts -> 64x{W}, 3 classes, batch 64 (8 distinct sweeps tiled 8x), HARD NMS",
                                    "ms_per_step": ms, "sweeps_per_s": 64 / (ms * 1e-3), "detections_per_step": ndet, "steps": steps}
        del hp

    # ---- batch-1 latency, the reference's protocol (tools/benchmark.py:231-238: sync, perf_counter, call, sync; warm-up 5) ----
    hp = HotPath("waymo", 1, dev, "HARD", fp_rate=args.fp_rate)

    def bench(fn):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3

    ms1 = hp.ms_of(hp.hd)
    for _ in range(5):
        hp.rasterize(); hp.dec.decode(ms1, hp.pp, hp.tasks, use_nms=True)
    t_r = [bench(hp.rasterize) for _ in range(20)]
    t_d = [bench(lambda: hp.dec.decode(ms1, hp.pp, hp.tasks, use_nms=True)) for _ in range(20)]
    g, det = capture(hp.step, dev)
    for _ in range(3):
        g.replay()
    (t,), _ = replay_timed([g], 20, flush)
    extra["batch1_latency"] = {"workload": "1 waymo-shaped sweep, HARD NMS", "protocol": "tools/benchmark.py:231-238 (sync-bracketed wall clock, warm-up 5)",
                               "rasterize_ms": float(np.mean(t_r)), "decoder_ms": float(np.mean(t_d)),
                               "total_ms": float(np.mean(t_r) + np.mean(t_d)), "fps": 1e3 / float(np.mean(t_r) + np.mean(t_d)),
                               "graph_replay_ms_device": float(np.mean(t)), "detections": int(det.wait())}
    return extra


def run_ours(args):
    import torch.distributed as dist
    from rv3d.distributed import PeerGather, gather_detections_fixed, pack_rows
    from rv3d import _native as N
    from rv3d._pipeline import run_nms

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the rv3d path has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    N.lib()

    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)   # before any pinned allocation: host buffers land next to the GPU's PCIe root

    B = args.batch
    n, H, W, C, M, ident = WORKLOADS[args.shape]
    gather_cap = B * C * PP["num_post_nms"]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    # the path's one exchange step (N > 1): detections go into every rank's gather buffer from inside the NMS pack
    # kernel (peer-memory stores over NVLink, per-slot sequence flags); NCCL all_gather only if symmetric memory is not
    # available on the box
    peer, gather_kind = None, "none (1 GPU)"
    if world > 1:
        try:
            peer = PeerGather(gather_cap, dev)
            gather_kind = "peer-memory stores fused into the pack kernel + per-slot sequence flags (no barrier kernel)"
        except Exception as exc:   # noqa: BLE001
            gather_kind = f"nccl all_gather_into_tensor (symmetric memory unavailable: {type(exc).__name__})"
    hp = HotPath(args.shape, B, dev, args.nms_mode, args.head_dtype, args.fp_rate, peer=peer, rank=rank)

    def full_step():
        det = hp.step(stats=hp.stats)
        if world > 1 and peer is None:
            m = det.params.shape[0]
            rows = torch.zeros((m, 13), dtype=torch.float32, device=dev)   # padded rows; counts travel in the block header
            rows[:, 0] = det.batch_index + float(rank * B); rows[:, 1] = det.categories; rows[:, 2] = det.scores; rows[:, 3:] = det.params
            gather_detections_fixed(rows[:gather_cap], gather_cap)
        return det

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # clocks are sampled from the warm-up to the end of the e2e loop (the device is under load throughout)
    clk = ClockSampler(local, enabled=(rank == 0))
    clk.__enter__()
    for _ in range(max(args.warmup, 3)):
        det = full_step()
    barrier()

    # ---------------- resident timing: the step as ONE CUDA graph, K replays, per-step events, L2 flush between ----------
    g_step, det = capture(full_step, dev)
    for _ in range(max(args.warmup, 3)):
        g_step.replay()
    barrier()
    hp.stats.zero_()
    (t_step,), wall = replay_timed([g_step], args.steps, flush, barrier if world > 1 else None)
    if peer is not None:
        peer.sync_steps()      # replays advance the device-side step counter; capturing advanced only the host's
    st = (hp.stats.cpu().numpy() / args.steps).tolist()
    ndet = det.wait()
    ncand = int(hp.dec._ws.get("counter", (1,), torch.int32, dev).item())
    total_ms = torch.tensor([t_step.sum()], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    value = world * B * args.steps / (total_ms * 1e-3)

    # ---------------- steady-state throughput: D step graphs in flight on D streams ----------------
    # One step leaves most of the machine idle at this batch size: its suppression kernel is one CTA per (sweep, class)
    # segment -- 48 CTAs on 148 SMs -- and latency-bound, while rasterize / decode are short bandwidth bursts.  A serving
    # loop therefore keeps several batches in flight.  D independent instances of the path (own inputs, workspaces and
    # gather buffers), each step still ONE graph replay through the public calls, replays issued round-robin on D
    # streams; device time from the first launch to the last completion.  No L2 flush here: a step's inputs (204 MB)
    # exceed the L2 (126 MB) and the D input sets rotate, so no step finds its inputs cached.
    D = max(1, args.pipeline_depth)
    pipelined = None
    if D > 1:
        lanes = []
        for d in range(D):
            pr = PeerGather(gather_cap, dev) if peer is not None else None
            h = HotPath(args.shape, B, dev, args.nms_mode, args.head_dtype, args.fp_rate, peer=pr, rank=rank,
                        inputs=(hp.sweeps, hp.head_host, hp.mapping))
            h.step(); h.step()
            barrier()
            g, det_d = capture(h.step, dev)
            if pr is not None:
                pr.sync_steps()
            lanes.append((h, torch.cuda.Stream(dev), g, det_d, pr))

        def run_lanes(K):
            main = torch.cuda.current_stream(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(main)
            for _, s_, _, _, _ in lanes:
                s_.wait_event(e0)
            for k in range(K):
                _, s_, g, _, _ = lanes[k % D]
                with torch.cuda.stream(s_):
                    g.replay()
            for _, s_, _, _, _ in lanes:
                main.wait_stream(s_)
            e1.record(main)
            return e0, e1

        run_lanes(max(args.warmup, 3) * D)
        barrier()
        t0 = time.perf_counter()
        e0, e1 = run_lanes(args.steps)
        barrier()
        wall_p = time.perf_counter() - t0
        ms_p = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms_p, op=dist.ReduceOp.MAX)
        ms_p = float(ms_p.item())
        for _, _, _, det_d, pr in lanes:
            assert det_d.wait() == ndet, "a pipelined lane produced a different number of detections"
            if pr is not None:
                pr.sync_steps()
        pipelined = {"depth": D, "ms_total": ms_p, "ms_per_step": ms_p / args.steps, "wall_ms_per_step": wall_p / args.steps * 1e3,
                     "value": world * B * args.steps / (ms_p * 1e-3)}
        del lanes

    # ---------------- N > 1: the gathered rows equal the concatenation of every rank's local output ----------------
    gather_ok = None
    if peer is not None:
        det = full_step()
        peer.wait_published()
        torch.cuda.synchronize()
        got = PeerGather.unpack_rows(peer.rows_published())
        m = det.wait()
        mine = pack_rows(det.params[:m], det.scores[:m], det.categories[:m], det.batch_index[:m], batch_offset=rank * B)
        from rv3d.distributed import gather_detections
        want = gather_detections(mine)
        ok = torch.tensor([int(got.shape == want.shape and torch.equal(got, want))], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        gather_ok = bool(ok.item())

    # ---------------- stage graphs: rasterize | decode_compact | bucketing + NMS + pack (roofline, breakdown) -------------
    hp1 = hp if peer is None else HotPath(args.shape, B, dev, args.nms_mode, args.head_dtype, args.fp_rate,
                                          inputs=(hp.sweeps, hp.head_host, hp.mapping))
    cand_box = {}

    def stage_decode():
        cand_box["c"] = hp1.dec.candidates(hp1.ms_of(hp1.hd), hp1.pp, hp1.tasks)

    def stage_nms():
        return run_nms(hp1.dec._ws, cand_box["c"], PP["num_pre_nms"], PP["num_post_nms"], PP["nms_threshold"], args.nms_mode,
                       N.OUT_QUAT, score_range=(PP["min_confidence"], 1.0))

    hp1.rasterize(); stage_decode(); stage_nms()
    torch.cuda.synchronize()
    g_r, _ = capture(hp1.rasterize, dev)
    g_d, _ = capture(stage_decode, dev)
    g_n, _ = capture(stage_nms, dev)

    def stage_raster_decode():
        hp1.rasterize(); stage_decode()

    def stage_raster_decode_forked():             # the pair the way a step runs it: rasterizer on a forked stream
        cur = torch.cuda.current_stream(dev)
        hp1.side.wait_stream(cur)
        with torch.cuda.stream(hp1.side):
            hp1.rasterize()
        stage_decode()
        cur.wait_stream(hp1.side)

    g_rd, _ = capture(stage_raster_decode, dev)   # the roofline pair as ONE graph: one launch, no gap between the two stages
    g_rdf, _ = capture(stage_raster_decode_forked, dev)
    for _ in range(3):
        g_r.replay(); g_d.replay(); g_n.replay(); g_rd.replay(); g_rdf.replay()
    (t_raster, t_decode, t_nms), _ = replay_timed([g_r, g_d, g_n], args.steps, flush)
    (t_rd,), _ = replay_timed([g_rd], args.steps, flush)
    (t_rdf,), _ = replay_timed([g_rdf], args.steps, flush)
    # the same pair on the same sweeps in a sensor's firing order (second record; the judged workload stays shuffled)
    from rv3d.math.range_view import pack_sweeps
    pts_keep = (hp1.pts, hp1.las, hp1.cnt)
    rep = B // len(hp.sweeps)
    fo = [t.to(dev) for t in pack_sweeps([firing_order(sw, W) for sw in hp.sweeps], dev)]
    hp1.pts, hp1.las, hp1.cnt = [t.repeat(rep, *([1] * (t.dim() - 1))) if rep > 1 else t for t in fo]
    hp1.rasterize(); torch.cuda.synchronize()
    g_rf, _ = capture(hp1.rasterize, dev)
    g_rdf_f, _ = capture(stage_raster_decode_forked, dev)
    for _ in range(3):
        g_rf.replay(); g_rdf_f.replay()
    (t_raster_f,), _ = replay_timed([g_rf], args.steps, flush)
    (t_rdf_f,), _ = replay_timed([g_rdf_f], args.steps, flush)
    hp1.pts, hp1.las, hp1.cnt = pts_keep
    del g_rf, g_rdf_f, fo

    # ---------------- e2e: pinned host inputs -> H2D -> path -> D2H of the detections, every step --------
    def e2e_leg(hpx, steps):
        copy_in, copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        bufs = [dict(pts=torch.empty_like(hpx.pts), las=torch.empty_like(hpx.las), cnt=torch.empty_like(hpx.cnt),
                     head={k: torch.empty_like(v) for k, v in hpx.hd.items()}, ready=torch.cuda.Event(), free=torch.cuda.Event())
                for _ in range(2)]
        h2d = (hpx.pts_h.numel() * 4 + hpx.las_h.numel() + hpx.cnt_h.numel() * 4 +
               sum(v.numel() * v.element_size() for v in hpx.head_h.values()))
        rows_per_rank = (gather_cap + 1) * 16 if peer is not None else gather_cap * 13 + 1
        out_h = [torch.empty((world if (peer is not None and rank == 0) else 1) * rows_per_rank, dtype=torch.float32).pin_memory()
                 for _ in range(2)]
        done = [torch.cuda.Event() for _ in range(2)]

        def upload(slot):
            b = bufs[slot]
            with torch.cuda.stream(copy_in):
                copy_in.wait_event(b["free"])
                b["pts"].copy_(hpx.pts_h, non_blocking=True); b["las"].copy_(hpx.las_h, non_blocking=True)
                b["cnt"].copy_(hpx.cnt_h, non_blocking=True)
                for k2, v in hpx.head_h.items():
                    b["head"][k2].copy_(v, non_blocking=True)
                b["ready"].record(copy_in)

        def run(nsteps):
            cur = torch.cuda.current_stream(dev)
            for b in bufs:
                b["free"].record(cur)
            d2h = 0
            upload(0)
            for k in range(nsteps):
                slot = k & 1
                if k + 1 < nsteps:
                    upload(slot ^ 1)          # next step's inputs cross PCIe while this step computes
                cur.wait_event(bufs[slot]["ready"])
                b = bufs[slot]
                # the calls a user makes: rv3d.math.range_view.rasterize_sweeps + RangeDecoder.decode_async
                det = hpx.step(b["pts"], b["las"], b["cnt"], b["head"])
                b["free"].record(cur)
                if peer is not None:
                    peer.wait_published()
                    # every rank holds the full gather on the device; the host copy is the whole set on rank 0 and the
                    # rank's own detections elsewhere (the reference's ranks each write only their own sweeps' files)
                    src = peer.rows_published().flatten() if rank == 0 else peer.rows_published()[rank].flatten()
                else:
                    src = det.buffer                     # padded detections + the count, one block
                done[slot].record(cur)
                with torch.cuda.stream(copy_out):        # the download overlaps the next step's kernels
                    copy_out.wait_event(done[slot])
                    out_h[slot][: src.numel()].copy_(src, non_blocking=True)
                    src.record_stream(copy_out)
                d2h = src.numel() * 4
            cur.wait_stream(copy_out)
            return d2h

        run(2)
        barrier()
        t0 = time.perf_counter()
        d2h = run(steps)
        barrier()
        secs = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(secs, op=dist.ReduceOp.MAX)
        secs = float(secs.item())
        return {"value": world * hpx.B * steps / secs, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "h2d_gbs_per_gpu": h2d * steps / secs / 1e9,
                "note": "host buffers pinned; upload double-buffered on a copy stream, download on a second one"}

    e2e = e2e_leg(hp, args.steps)
    e2e_f16 = None
    if args.head_dtype == "f32":
        hp16 = HotPath(args.shape, B, dev, args.nms_mode, "f16", args.fp_rate, peer=peer, rank=rank,
                       inputs=(hp.sweeps, hp.head_host, hp.mapping))
        hp16.step(); hp16.step()
        e2e_f16 = e2e_leg(hp16, args.steps)
        e2e_f16["note"] = ("second record: logits / regressands in float16 (the reference's own eval_precision 16 operating point, "
                           "detector.py:329-333), cart float32; " + e2e_f16["note"])
        del hp16
    clk.__exit__()
    clocks = clk.summary()

    extra = {}
    if world > 1 and args.global_batch and peer is not None:
        # BASELINE config 5: a fixed global batch sharded over the ranks (strong scaling of one big batch)
        Bg = args.global_batch // world
        peer5 = PeerGather(Bg * C * PP["num_post_nms"], dev)
        hp5 = HotPath(args.shape, Bg, dev, args.nms_mode, args.head_dtype, args.fp_rate, gen_batch=B, peer=peer5, rank=rank,
                      inputs=(hp.sweeps, hp.head_host, hp.mapping))
        hp5.step(); hp5.step()
        barrier()
        g5, _ = capture(hp5.step, dev)
        peer5.sync_steps()
        for _ in range(2):
            g5.replay()
        (t5,), _ = replay_timed([g5], 5, flush, barrier)
        t5 = torch.tensor([t5.sum()], dtype=torch.float64, device=dev)
        dist.all_reduce(t5, op=dist.ReduceOp.MAX)
        extra["config5_global_batch"] = {"global_batch": Bg * world, "batch_per_gpu": Bg, "steps": 5,
                                         "ms_per_step": float(t5.item()) / 5, "sweeps_per_s": Bg * world * 5 / (float(t5.item()) * 1e-3),
                                         "note": f"waymo shape, {B} distinct sweeps tiled to {Bg} per rank, detections gathered by the fused peer stores"}
        del hp5, peer5
    if world == 1 and not args.no_extras:
        extra.update(run_extras(args, dev, flush, all_cpus))

    if rank == 0:
        raster_b, decode_b = algorithmic_bytes(args.shape, B, ncand, 2 if args.head_dtype == "f16" else 4)
        peaks = {}
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
        except (OSError, ValueError):
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        rd_ms = float(np.mean(t_rdf))          # the pair the way a step runs it (rasterizer forked); serial form beside it
        achieved = (raster_b + decode_b) / (rd_ms * 1e-3) / 1e9
        gbs = lambda ms: (raster_b + decode_b) / (ms * 1e-3) / 1e9   # noqa: E731
        S = B * C
        single = {"value": value, "ms_per_step": total_ms / args.steps, "ms_per_step_median_rank0": float(np.median(t_step)),
                  "ms_per_step_max_rank0": float(np.max(t_step)),
                  "note": "one step at a time on one stream (the step's latency), 256 MiB L2 flush between steps"}
        if pipelined is not None:
            value, ms_step = pipelined["value"], pipelined["ms_per_step"]
            timed = (f"{args.steps} steps, each ONE CUDA-graph replay of rasterize_sweeps (forked stream inside the graph) + RangeDecoder.decode_async (no host read inside a step), "
                     f"issued round-robin on {D} streams over {D} independent input sets; CUDA events from the first launch to the last completion")
            l2 = "no flush: a step's inputs (204 MB) exceed the 126 MB L2 and the input sets rotate; single-stream latency figures flush 256 MiB between steps"
        else:
            ms_step = total_ms / args.steps
            timed = "one CUDA-graph replay of rasterize_sweeps (forked stream) + RangeDecoder.decode_async per step (no host read inside the step)"
            l2 = "256 MiB L2 flush between timed steps (outside the per-step CUDA events)"
        line = {
            "metric": METRIC, "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step,
            "single_stream": single, "pipeline_depth": D, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, B), detection_gather=gather_kind,
                           per_rank_data="identical synthetic sweeps on every rank (seed 1000)", host_affinity=numa,
                           head_dtype=args.head_dtype, timed_region=timed, l2=l2),
            "e2e": e2e,
            # own kernels per step: raster scatter + resolve, decode_compact, hist, bin_scan, scatter_records, nms_pull, pack
            # (+ kept_scan above 512 segments, + the peer wait at N > 1); memsets are not counted
            "gpu_launches": (8 + (1 if S > 512 else 0) + (1 if peer is not None else 0)) * args.steps,
            "pipelined": pipelined,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernels": "rasterize (scatter+resolve) + decode_compact", "achieved": achieved,
                         "peak": peak, "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(args.shape, B),
                         "algorithmic_bytes": {"rasterize": raster_b, "decode": decode_b},
                         "rasterize_gbs": raster_b / (float(np.mean(t_raster)) * 1e-3) / 1e9,
                         "decode_gbs": decode_b / (float(np.mean(t_decode)) * 1e-3) / 1e9,
                         "timing": ("achieved / frac: rasterize + decode_compact replayed as ONE CUDA graph the way a step runs the pair "
                                    "(the rasterizer's memset / scatter / resolve on a forked stream next to the counter fill + decode_compact, "
                                    "joined at the end), CUDA events on the replay stream around the replay, L2 flushed before every "
                                    "replay; serial: the same kernels back to back on one stream; rasterize_gbs / decode_gbs: each stage as "
                                    "its own graph (each pays its own graph launch)"),
                         "ms": rd_ms,
                         "serial": {"ms": float(np.mean(t_rd)), "achieved": gbs(float(np.mean(t_rd))), "frac": gbs(float(np.mean(t_rd))) / peak,
                                    "note": "scatter -> resolve -> decode_compact back to back on one stream inside one graph"},
                         # second record: the same sweeps in a sensor's firing order (the judged workload above is shuffled)
                         "firing_order": {"ms": float(np.mean(t_rdf_f)), "achieved": gbs(float(np.mean(t_rdf_f))),
                                          "frac": gbs(float(np.mean(t_rdf_f))) / peak,
                                          "rasterize_ms": float(np.mean(t_raster_f)),
                                          "rasterize_gbs": raster_b / (float(np.mean(t_raster_f)) * 1e-3) / 1e9,
                                          "note": ("same points sorted azimuth step by azimuth step like a spinning lidar emits them: the "
                                                   "scattered atomics and gathers of the rasterizer coalesce")},
                         # the suppression stage, for completeness: it reads each candidate's key + box once and writes the
                         # detections (NMS is latency / issue bound, an HBM fraction says little -- work units under "nms")
                         "nms_stage": {"bound": "hbm", "algorithmic_bytes": int(ncand) * 40 + int(ndet) * 52,
                                       "achieved": (int(ncand) * 40 + int(ndet) * 52) / (float(np.mean(t_nms)) * 1e-3) / 1e9,
                                       "frac": (int(ncand) * 40 + int(ndet) * 52) / (float(np.mean(t_nms)) * 1e-3) / 1e9 / peak,
                                       "note": "score bucketing + nms_pull_kernel + pack; not HBM-bound"}},
            "stage_ms": {"rasterize": float(np.mean(t_raster)), "decode_compact": float(np.mean(t_decode)),
                         "bucketing+nms+pack": float(np.mean(t_nms)), "wall_per_step_incl_flush": wall * 1e3,
                         "flush_ms_note": "wall includes the 256 MiB flush memset (~0.04 ms); host launch cost per step = one graph launch"},
            # work units of the suppression stage (device counters of nms_pull_kernel, averaged over the timed steps)
            "nms": {"candidates_per_step": int(ncand), "detections_per_step": int(ndet), "segments": S,
                    "candidates_consumed_per_step": st[20], "exact_iou_per_step": st[0], "approx_iou_per_step": st[19],
                    "kept_per_step": st[1], "frontier_rounds_per_step": st[2], "circle_tests_per_step": st[3],
                    "pairs_above_thr_per_step": st[18],
                    # cycles summed over the segments' CTAs per phase: window sort, pull walk, pull IoU, frontier pairs, greedy, publish
                    "phase_mcycles_per_step": [round(x / 1e6, 3) for x in st[4:10]],
                    "sub_phase_mcycles_per_step": {"frontier_walk": round(st[21] / 1e6, 3), "frontier_eval": round(st[22] / 1e6, 3),
                                                   "pull_walk_scan": round(st[23] / 1e6, 3)},
                    "slowest_segment_mcycles": round(float(hp.stats[10].item()) / 1e6, 3),
                    "largest_segment": int(hp.stats[11].item()),
                    "slowest_segment_phase_kcycles": [int(x) // 1000 for x in hp.stats[12:18].tolist()]},
        }
        if e2e_f16 is not None:
            line["e2e_f16_heads"] = e2e_f16
        if gather_ok is not None:
            line["gather_ok"] = gather_ok
        if world == 1 and not args.no_cpu_baseline and args.head_dtype == "f32":
            import oracle  # noqa: F401
            os.sched_setaffinity(0, all_cpus)          # the CPU baseline gets every host core again
            cores = len(all_cpus)
            torch.set_num_threads(cores)
            pool = ThreadPoolExecutor(max_workers=cores)
            h1 = {k: v[:1] for k, v in hp.head_host.items()}
            cpu_step(hp.sweeps[0], h1, hp.mapping, args.shape, args.nms_mode, pool)        # JIT / page-in
            rs = [cpu_step(hp.sweeps[0], h1, hp.mapping, args.shape, args.nms_mode, pool) for _ in range(3)]
            line["cpu_baseline"] = cpu_baseline_record(rs, cores, args.shape)
        if extra:
            line["extra"] = extra
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shape", default="waymo", choices=["waymo", "av2"])
    ap.add_argument("--batch", type=int, default=16, help="sweeps per GPU per step")
    ap.add_argument("--nms-mode", default="HARD", choices=["HARD", "WEIGHTED"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the BASELINE config 1 / 3 / 4 and batch-1 latency legs")
    ap.add_argument("--pipeline-depth", type=int, default=6,
                    help="step graphs kept in flight on as many streams for the throughput figure (1 = one step at a time)")
    ap.add_argument("--global-batch", type=int, default=512,
                    help="N > 1: also time BASELINE config 5, this many sweeps sharded over the ranks (0 = skip)")
    ap.add_argument("--head-dtype", default="f32", choices=["f32", "f16"],
                    help="exploration only: f16 = half-precision logits / regressands next to float32 cart (autocast); "
                         "the CPU baseline leg is skipped")
    ap.add_argument("--fp-rate", type=float, default=None,
                    help="exploration only: fraction of background pixels that fire (default: the SURVEY 8d density, 0.95)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
