# round-2 evidence: launch list of the bench command + one --set full capture of every own kernel of a step
TAG=${1:-r02}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_scatter|raster_resolve|decode_compact|nms_pull|pack_kernel|hist_kernel|bin_scan|scatter_records" -c 9 -o gpurun_out/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_full.log 2>&1
tail -2 gpurun_out/${TAG}_full.log | cut -c1-200
python tools/ncu_summary.py gpurun_out/${TAG}_full.ncu-rep > gpurun_out/${TAG}_ncu_full.md
cat gpurun_out/${TAG}_ncu_full.md
