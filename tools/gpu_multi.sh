N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/check_peer_gather.py 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 1500 gpurun_out/bench_n$N.err
python - <<EOF
import json
d=json.loads(open("gpurun_out/bench_n$N.json").read().strip().splitlines()[-1])
print("N=$N", d["value"], d["ms_per_step"], d["stage_ms"], "gather_ok", d.get("gather_ok"), "e2e", d["e2e"]["value"], d.get("e2e_f16_heads",{}).get("value"))
print(d.get("extra"))
print(d["config"]["detection_gather"])
EOF
