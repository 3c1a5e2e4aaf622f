timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --pipeline-depth 1 > gpurun_out/r02b_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/r02b_launches.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:70]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
for k, (n, t) in agg.items():
    print(f"{t/n/1000:9.2f} us avg  x{n:5d}  {k}")
PY
