# dev round: NMS parity tests + a short bench (no extras), key figures
[ -n "$SKIP_TESTS" ] || timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py -x -q -m gpu 2>&1 | tail -4
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_k.json 2> gpurun_out/bench_k.err; tail -5 gpurun_out/bench_k.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_k.json'))
print('value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'single', round(d['single_stream']['ms_per_step'], 4), 'stages', {k: round(v, 4) for k, v in d['stage_ms'].items() if isinstance(v, float)})
print('roofline', round(d['roofline']['frac'], 4), d['roofline']['ms'], 'nms', {k: d['nms'][k] for k in ("phase_mcycles_per_step", "sub_phase_mcycles_per_step", "slowest_segment_phase_kcycles")})
PY
