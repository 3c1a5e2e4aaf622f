# dev: whole-stage time against the NMS cluster size (RV3D_NMS_CLUSTER overrides the library's choice)
run() { python bench.py --steps 8 --warmup 3 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('   ', round(d['value']), d['stage_ms']['sort+nms+pack'])"; }
for cfg in "--batch 32" "--batch 16 --nms-mode WEIGHTED" "--batch 4 --shape av2" "--batch 8"; do
  for P in 0 1 2 3 4 8; do echo "$cfg P=$P"; RV3D_NMS_CLUSTER=$P run $cfg; done
done
