timeout 1200 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py tests/test_gpu_iou_decisions.py tests/test_gpu_assign.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; tail -5 gpurun_out/bench_j.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_j.json'))
print('HARD value', d['value'], 'ms', d['ms_per_step'], 'single', d['single_stream']['ms_per_step'], d['stage_ms']['bucketing+nms+pack'])
for k in ('config1_av2_hard','config3_stress_200k_weighted','config3_stress_200k_hard','config4_w900','config4_w3600','batch1_latency'):
    v = d['extra'][k]; print(k, v.get('ms_per_call') or v.get('ms_per_step') or v.get('graph_replay_ms_device'))
PY
timeout 300 python tools/bench_8f.py 2>/dev/null | grep -E "box_iou|BEV" | python -c "
import sys, json
for l in sys.stdin:
    d=json.loads(l); print(d['kernel'], d['ms'])
"
