# dev round: the rotated-IoU bit-equality tests + NMS parity + a short bench
timeout 600 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py tests/test_gpu_iou_decisions.py tests/test_gpu_assign.py -x -q -m gpu 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_m.json 2> gpurun_out/bench_m.err; tail -3 gpurun_out/bench_m.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_m.json'))
print('value', round(d['value']), 'single', round(d['single_stream']['ms_per_step'], 4), {k: round(v, 4) for k, v in d['stage_ms'].items() if isinstance(v, float)})
print('nms', d['nms']['phase_mcycles_per_step'], d['nms']['sub_phase_mcycles_per_step'])
PY
