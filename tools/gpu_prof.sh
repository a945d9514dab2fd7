#!/bin/bash
# GPU visit: parity tests, bench line, ncu --set full of selected kernels.  Usage: bash tools/gpu_prof.sh <tag> <kernel-regex> [count]
set -u
TAG=${1:-p}
RE=${2:-raster_bwd_kernel|raster_fwd_kernel}
CNT=${3:-2}
OUT=gpurun_out
mkdir -p $OUT
bash tools/gpu_quick.sh $TAG
echo "== ncu --set full ($RE)"
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $((CNT*3)) -c $CNT \
    -o $OUT/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_full.ncu-rep
