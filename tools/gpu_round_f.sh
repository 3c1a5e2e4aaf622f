timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py -x -q -m gpu 2>&1 | tail -5
for t in 0; do
RV3D_NMS_NO_TAIL=$t timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nms-mode WEIGHTED --pipeline-depth 1 > gpurun_out/bench_gw.json 2> gpurun_out/bench_gw.err; tail -5 gpurun_out/bench_gw.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_gw.json'))
print('NO_TAIL=$t WEIGHTED value', d['value'], 'ms', d['ms_per_step'], d['stage_ms']['bucketing+nms+pack'], d['nms'])
for k in ('config3_stress_200k_weighted','config3_stress_200k_hard'): print(k, d['extra'][k]['ms_per_call'])
PY
done
