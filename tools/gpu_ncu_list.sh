#!/bin/bash
# ncu launch list of the bench command (per-launch durations; cold-cache, serialised) + one --set full capture of raster_bwd.
set -u
TAG=${1:-x}
OUT=gpurun_out
mkdir -p $OUT
timeout 75 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
wc -l $OUT/${TAG}_launches.csv
timeout 45 ncu --set full --clock-control none -k regex:raster_bwd_kernel -s 3 -c 1 -o $OUT/${TAG}_raster_bwd_full \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_raster_bwd_full.ncu-rep 2>&1
echo done
