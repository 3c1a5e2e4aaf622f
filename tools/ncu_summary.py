"""Summarise an .ncu-rep (ncu --set full) into a small markdown table of the metrics DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--traffic profiles/traffic.json] > profiles/<name>.md

--traffic: also write dram__bytes_read.sum + dram__bytes_write.sum of the FIRST launch of raster_scatter, raster_resolve and
decode_compact (the roofline pair of bench.py) as {"shape", "batch", "bytes_per_step", "per_kernel", "source"}.
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM % of peak"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed.sum", "thread instructions"),
    ("sm__inst_executed_pipe_fp64.sum", "fp64 pipe instr"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("lts__t_bytes.sum", "L2 bytes"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {name: hdr.index(name) for name, _ in WANT if name in hdr}
    kn = hdr.index("Kernel Name")
    print(f"# ncu --set full summary of `{path.split('/')[-1]}`\n")
    print("| kernel | " + " | ".join(label for name, label in WANT if name in idx) + " |")
    print("|---|" + "---|" * len(idx))
    for r in rows[2:]:
        cells = []
        for name, _ in WANT:
            if name in idx:
                cells.append(f"{r[idx[name]]} {units[idx[name]]}".strip())
        print(f"| `{r[kn][:60]}` | " + " | ".join(cells) + " |")


def traffic(path, out):
    import json
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per = {}
    for r in rows[2:]:
        for name in ("raster_scatter", "raster_resolve", "decode_compact"):
            if name in r[kn] and name not in per:
                per[name] = int(float(r[rd].replace(",", "")) * scale[units[rd]] + float(r[wr].replace(",", "")) * scale[units[wr]])
    json.dump({"shape": "waymo", "batch": 16, "bytes_per_step": sum(per.values()), "per_kernel": per,
               "source": f"ncu --set full, {path.split('/')[-1]} (cold cache per kernel: ncu flushes L2 between kernels, so part of "
                         "the stores is still in L2 when a kernel ends and is not counted)"}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1])
    if "--traffic" in sys.argv:
        traffic(sys.argv[1], sys.argv[sys.argv.index("--traffic") + 1])
