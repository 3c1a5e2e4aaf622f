# full GPU suite + the complete bench line (extras, CPU baseline) + 8f bench
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -3 gpurun_out/r02_bench_line.err
python - <<PY
import json
d = json.load(open('gpurun_out/r02_bench_line.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'single', round(d['single_stream']['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'], 4))
print({k: round(v, 4) for k, v in d['stage_ms'].items() if isinstance(v, float)})
for k, v in d.get('extra', {}).items():
    if isinstance(v, dict):
        print(k, v.get('ms_per_call') or v.get('ms_per_step') or v.get('graph_replay_ms_device') or v.get('value'))
PY
