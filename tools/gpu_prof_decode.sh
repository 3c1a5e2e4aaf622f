TAG=${1:-r02e}
timeout 600 ncu --metrics gpu__time_duration.sum --cache-control none --clock-control none -k regex:"decode_compact|raster_" -c 40 python tools/dev_time_stages.py 2>&1 | grep -E "^\s+(void )?(rv3d::)?(decode_compact|raster_)|gpu__time" | paste - - | awk '{print $1,$2, $(NF)}' | sort | uniq -c | sort -rn | head -12
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"decode_compact" -c 2 -o gpurun_out/prof_$TAG python tools/dev_time_stages.py > gpurun_out/prof_$TAG.log 2>&1
python tools/ncu_summary.py gpurun_out/prof_$TAG.ncu-rep
