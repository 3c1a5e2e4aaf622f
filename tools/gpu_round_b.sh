timeout 900 python -m pytest tests/test_gpu_assign.py tests/test_gpu_prep.py -x -q -m gpu 2>&1 | tail -5
timeout 600 python tools/bench_8f.py > gpurun_out/bench_8f.jsonl 2> gpurun_out/bench_8f.err; tail -5 gpurun_out/bench_8f.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_8f.jsonl'):
    d = json.loads(l); print(f"{d['ms']*1000:9.1f} us  frac {d['frac']:.3f}  {d['kernel']}")
PY
bash tools/gpu_prof_8f.sh 2>&1 | grep -E "rv3d::|Memset|memset" | head -30
