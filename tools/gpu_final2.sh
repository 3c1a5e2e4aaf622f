# final evidence of the round: full GPU suite, full bench line, WEIGHTED line, ncu launch list + --set full summary
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -3 gpurun_out/r02_bench_line.err
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras --nms-mode WEIGHTED > gpurun_out/r02_bench_weighted.json 2> gpurun_out/r02_bench_weighted.err; tail -3 gpurun_out/r02_bench_weighted.err
python - <<PY
import json
for f in ('r02_bench_line', 'r02_bench_weighted'):
    d = json.load(open(f'gpurun_out/{f}.json'))
    print(f, 'value', round(d['value']), 'ms/step', round(d['ms_per_step'], 4), 'single', round(d['single_stream']['ms_per_step'], 4), 'e2e', round(d['e2e']['value']), 'roofline', round(d['roofline']['frac'], 4), {k: round(v, 4) for k, v in d['stage_ms'].items() if isinstance(v, float)})
PY
bash tools/gpu_prof_final.sh r02f 2>&1 | tail -14
