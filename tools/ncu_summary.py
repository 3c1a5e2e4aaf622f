"""Summarise an .ncu-rep (ncu --set full) into a small markdown table of the metrics DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/<name>.md
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


if __name__ == "__main__":
    main(sys.argv[1])
