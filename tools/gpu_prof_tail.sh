timeout 900 ncu --set full --clock-control none --import-source on -k regex:"wnms_tail" -c 1 -o gpurun_out/tail python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras --nms-mode WEIGHTED --pipeline-depth 1 > gpurun_out/tail.log 2>&1
tail -3 gpurun_out/tail.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/tail.ncu-rep 2>&1 | tail -5
