"""Warp-instruction share of source-line ranges from an `ncu --page source --csv` dump.
    python tools/ncu_regions.py src.csv name:file:first:last ...
"""
import csv, sys
from collections import defaultdict

def load(path):
    rows = list(csv.reader(open(path)))
    agg = defaultdict(int); cur = None; i = 0
    while i < len(rows):
        r = rows[i]
        if r and r[0] in ("File Name", "File Path"): cur = r[1]
        elif r and r[0] == "Line No" and len(r) > 8:
            ie = r.index("Instructions Executed"); j = i + 1
            while j < len(rows) and rows[j] and rows[j][0] not in ("File Name", "File Path", "Line No"):
                q = rows[j]
                if q[0].strip() and len(q) > ie:
                    try: agg[((cur or "?").split('/')[-1], int(q[0]))] += int(q[ie] or 0)
                    except ValueError: pass
                j += 1
            i = j - 1
        i += 1
    return agg

if __name__ == "__main__":
    agg = load(sys.argv[1]); tot = sum(agg.values())
    print("total", tot)
    files = defaultdict(int)
    for (f, l), v in agg.items(): files[f] += v
    print({k: round(v / tot * 100, 1) for k, v in files.items()})
    for spec in sys.argv[2:]:
        name, f, a, b = spec.split(":")
        s = sum(v for (ff, l), v in agg.items() if ff == f and int(a) <= l <= int(b))
        print(f"{name:24s} {s / tot * 100:5.1f}%  {s}")
