for d in ${DEPTHS:-3 4 5 6 8}; do
  timeout 300 python bench.py --steps 24 --warmup 3 --no-extras --no-cpu-baseline --pipeline-depth $d > gpurun_out/pd.json 2> gpurun_out/pd.err
  tail -c 600 gpurun_out/pd.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/pd.json"))
print("depth $d:", round(d["value"]), d["ms_per_step"], d["pipelined"], "single", round(d["single_stream"]["value"]))
EOF
done
