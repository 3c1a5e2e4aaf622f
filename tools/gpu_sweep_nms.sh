IFS=";" read -ra LIST <<< "${CFGS:-2 3;2 4;1.5 2.5;1.5 3;2.5 3.5;2 2.5}"
for cfg in "${LIST[@]}"; do
  IFS=" " read -r a b <<< "$cfg"; set -- $a $b
  RV3D_NMS_RCAP=$1 RV3D_NMS_CELL=$2 timeout 200 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - <<EOF
import json
d=json.load(open("gpurun_out/sw.json"))
print("rcap $1 cell $2:", round(d["value"]), round(d["single_stream"]["ms_per_step"],4), d["stage_ms"]["bucketing+nms+pack"], d["nms"]["phase_mcycles_per_step"], d["nms"]["slowest_segment_mcycles"])
EOF
done
