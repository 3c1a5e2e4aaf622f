timeout 900 python -m pytest tests/test_gpu_nms.py tests/test_gpu_bench_parity.py -x -q 2>&1 | tail -8
timeout 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err
tail -c 1500 gpurun_out/bench_c.err
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_c.json"))
print(d["value"], d["ms_per_step"], d["stage_ms"]); print(d["nms"]); print(d["roofline"]["frac"], d["e2e"]["value"])
EOF
