# dev round: short bench WITH extras, key figures
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_l.json 2> gpurun_out/bench_l.err; tail -5 gpurun_out/bench_l.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_l.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'single', round(d['single_stream']['ms_per_step'], 4), {k: round(v, 4) for k, v in d['stage_ms'].items() if isinstance(v, float)})
print('nms', d['nms']['phase_mcycles_per_step'], d['nms']['sub_phase_mcycles_per_step'], d['nms']['frontier_rounds_per_step'])
for k, v in d.get('extra', {}).items():
    if isinstance(v, dict):
        print(k, v.get('ms_per_call') or v.get('ms_per_step') or v.get('graph_replay_ms_device') or v.get('value'))
PY
