timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py tests/test_gpu_decode.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_g.json 2> gpurun_out/bench_g.err; tail -5 gpurun_out/bench_g.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --nms-mode WEIGHTED > gpurun_out/bench_gw.json 2> gpurun_out/bench_gw.err; tail -5 gpurun_out/bench_gw.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_g.json'))
print('HARD value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'])
for k in ('config3_stress_200k_weighted','config3_stress_200k_hard'): print(k, d['extra'][k]['ms_per_call'])
d = json.load(open('gpurun_out/bench_gw.json'))
print('WEIGHTED value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'], d['stage_ms'])
PY
