timeout 900 python -m pytest tests/test_gpu_assign.py tests/test_gpu_prep.py -x -q -m gpu 2>&1 | tail -15
timeout 600 python tools/bench_8f.py > gpurun_out/bench_8f.jsonl 2> gpurun_out/bench_8f.err; tail -5 gpurun_out/bench_8f.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_8f.jsonl'):
    d = json.loads(l); print(f"{d['ms']*1000:9.1f} us  frac {d['frac']:.3f}  {d['kernel']}")
PY
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/bench_d.json 2> gpurun_out/bench_d.err; tail -5 gpurun_out/bench_d.err
python - <<'PY'
import json
d = json.load(open('gpurun_out/bench_d.json'))
print('value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'])
r = d['roofline']; print('frac', r['frac'], 'ms', r['ms'], 'serial', r['serial'], 'firing', r['firing_order'])
print(d['stage_ms'])
PY
