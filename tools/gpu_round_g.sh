timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py tests/test_gpu_decode.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/bench_h.json 2> gpurun_out/bench_h.err; tail -5 gpurun_out/bench_h.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_h.json'))
print('HARD value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'], d['stage_ms'])
PY
