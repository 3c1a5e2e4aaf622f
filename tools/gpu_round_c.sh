timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py -x -q -m gpu 2>&1 | tail -5
for d in 6 8; do
timeout 600 python bench.py --steps 24 --warmup 5 --no-extras --no-cpu-baseline --pipeline-depth $d > gpurun_out/bench_e$d.json 2> gpurun_out/bench_e$d.err; tail -5 gpurun_out/bench_e$d.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_e$d.json'))
print('depth $d value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'])
r = d['roofline']; print('frac', r['frac'], 'ms', r['ms'], 'serial', r['serial']['ms'])
print(d['stage_ms'])
PY
done
