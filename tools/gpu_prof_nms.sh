# ncu --set full capture of the NMS kernel (source-level)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nms_pull -c 1 -o gpurun_out/r02_nms python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_nms.log 2>&1
tail -2 gpurun_out/ncu_nms.log | cut -c1-300
ncu -i gpurun_out/r02_nms.ncu-rep --page source --csv --print-source cuda,sass -k regex:nms_pull --launch-count 1 > gpurun_out/r02_nms_src.csv 2>/dev/null
python tools/ncu_hot_lines.py gpurun_out/r02_nms_src.csv "" 60 2>&1 | tail -62
