TAG=${1:-r02b}
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -c 60 --csv --log-file gpurun_out/warm_$TAG.csv python tools/dev_time_raster.py > /dev/null 2>&1
python - <<PY
import csv,collections
rows=list(csv.reader(open("gpurun_out/warm_$TAG.csv")))
hdr=[i for i,r in enumerate(rows) if "Kernel Name" in r][0]
h=rows[hdr]; kn=h.index("Kernel Name"); mv=h.index("Metric Value")
d=collections.defaultdict(list)
for r in rows[hdr+1:]:
    if len(r)>mv: d[r[kn][:50]].append(float(r[mv].replace(",","")))
for k,v in d.items(): print(k, len(v), "median ns", sorted(v)[len(v)//2])
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_scatter|raster_resolve" -c 4 -o gpurun_out/prof_$TAG python tools/dev_time_raster.py > gpurun_out/prof_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep
