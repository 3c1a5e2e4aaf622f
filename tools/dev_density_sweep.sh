# dev: the bench at other candidate densities / shapes / batch sizes (exploration; the judged line is the default run)
show() { tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(sys.argv[1], round(d['value']), 'sweeps/s', {k: round(v, 4) for k, v in d['stage_ms'].items()}, 'frac', round(d['roofline']['frac'], 3), 'candidates', d['nms']['candidates_per_step'], 'detections', d['nms']['detections_per_step'])" "$1"; }
for fp in 0.05 0.3; do python bench.py --steps 10 --warmup 3 --no-cpu-baseline --fp-rate $fp 2>&1 | show "fp_rate=$fp"; done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --shape av2 2>&1 | show "av2"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 64 2>&1 | show "batch=64"
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --nms-mode WEIGHTED 2>&1 | show "weighted"
