timeout 900 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -30
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b.json 2> gpurun_out/bench_b.err
tail -c 2000 gpurun_out/bench_b.err
python - <<EOF
import json
d=json.load(open("gpurun_out/bench_b.json"))
print(d["value"], d["ms_per_step"], d["stage_ms"], d["nms"], d["e2e"]["value"])
print({k:(v.get("ms_per_step") or v.get("ms_per_call") or v) for k,v in d.get("extra",{}).items()})
EOF
