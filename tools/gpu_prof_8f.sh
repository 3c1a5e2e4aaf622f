# launch list (kernel durations) of the 8f bench: which kernels / glue each row spends its time in
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/8f_launches.csv python tools/bench_8f.py > gpurun_out/8f_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(l for l in open('gpurun_out/8f_launches.csv') if l.startswith('"')))
h = rows[0]; ki = h.index('Kernel Name'); vi = h.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki][:90]
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(',', ''))
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:60]:
    print(f"{t/n/1000:9.2f} us avg  x{n:5d}  {k}")
PY
