# ncu: launch list of a whole step + --set full capture (source-level) of rasterize / decode kernels
TAG=${1:-r02a}
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/launches_$TAG.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"raster_scatter|raster_resolve|decode_compact" -c 6 -o gpurun_out/prof_$TAG python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/prof_$TAG.log 2>&1
tail -2 gpurun_out/prof_$TAG.log | cut -c1-300
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep
