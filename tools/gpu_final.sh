# final evidence of the round: default bench line, launch list, --set full capture, 8f bench
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_line.json 2> gpurun_out/r02_bench_line.err; tail -3 gpurun_out/r02_bench_line.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; tail -3 gpurun_out/r02_bench_reference.err
timeout 600 python tools/bench_8f.py > gpurun_out/r02_bench_8f.jsonl 2> gpurun_out/bench_8f.err; tail -3 gpurun_out/bench_8f.err
bash tools/gpu_prof_final.sh r02c 2>&1 | tail -14
python -c "
import json
d=json.load(open('gpurun_out/r02_bench_line.json'))
print(d['value'], d['ms_per_step'], d['single_stream']['ms_per_step'], d['roofline']['frac'], d['roofline']['serial']['frac'], d['e2e']['value'], d['e2e_f16_heads']['value'])
print(open('gpurun_out/r02_bench_reference.json').read()[:600])
"
