"""Aggregate an `ncu --page source --csv --print-source cuda,sass` dump per CUDA source line.

    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass -k regex:NAME --launch-count 1 > src.csv
    python tools/ncu_hot_lines.py src.csv [file-substring] [top]
"""
import csv
import sys
from collections import defaultdict


def main(path, want="", top=40):
    rows = list(csv.reader(open(path)))
    cur_file = None
    agg = defaultdict(lambda: [0, 0, ""])
    i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] in ("File Name", "File Path"):
            cur_file = r[1]
        elif r and r[0] == "Line No" and len(r) > 8:
            hdr = r
            ie = hdr.index("Instructions Executed")
            ws = hdr.index("Warp Stall Sampling (All Samples)")
            j = i + 1
            line = None
            while j < len(rows) and rows[j] and rows[j][0] not in ("File Name", "File Path", "Line No"):
                q = rows[j]
                if q[0].strip() and len(q) > ie:   # a CUDA source line (its SASS rows follow, line no empty)
                    line = int(q[0])
                    agg[(cur_file, line)][2] = q[1]
                    try:
                        agg[(cur_file, line)][0] += int(q[ie] or 0)
                        agg[(cur_file, line)][1] += int(q[ws] or 0)
                    except ValueError:
                        pass
                j += 1
            i = j - 1
        i += 1
    tot = sum(v[0] for v in agg.values()) or 1
    tots = sum(v[1] for v in agg.values()) or 1
    print(f"total warp instructions {tot}, stall samples {tots}")
    items = [(k, v) for k, v in agg.items() if want in (k[0] or "")]
    print("-- by instructions")
    for (f, ln), v in sorted(items, key=lambda kv: -kv[1][0])[:top]:
        print(f"{v[0]/tot*100:5.1f}% i {v[1]/tots*100:5.1f}% s  {(f or "?").split('/')[-1]}:{ln}: {v[2].strip()[:100]}")
    print("-- by stall samples")
    for (f, ln), v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        print(f"{v[0]/tot*100:5.1f}% i {v[1]/tots*100:5.1f}% s  {(f or "?").split('/')[-1]}:{ln}: {v[2].strip()[:100]}")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else "", int(sys.argv[3]) if len(sys.argv) > 3 else 40)
