timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nms-mode WEIGHTED --pipeline-depth 1 > gpurun_out/bench_gw.json 2> gpurun_out/bench_gw.err; tail -5 gpurun_out/bench_gw.err
python - <<PY
import json
d = json.load(open('gpurun_out/bench_gw.json'))
print('WEIGHTED value', d['value'], 'ms', d['ms_per_step'], d['stage_ms']['bucketing+nms+pack'])
for k in ('config1_av2_hard','config3_stress_200k_weighted','config3_stress_200k_hard','config4_w900','config4_w3600','batch1_latency'):
    v = d['extra'][k]; print(k, v.get('ms_per_call') or v.get('ms_per_step') or v.get('graph_replay_ms_device'))
PY
