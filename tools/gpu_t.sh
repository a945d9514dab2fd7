#!/bin/bash
# Very short GPU visit: parity tests of the 8f-4 components, their timing, one ncu --set full capture of their kernels.
# Usage (under gpurun): bash tools/gpu_t.sh <tag> [extra pytest files]
set -u
TAG=${1:-t}
shift
OUT=gpurun_out
mkdir -p $OUT
echo "== parity tests"
timeout 240 python -m pytest tests/test_gpu_z_next_voxel_lbs.py tests/test_gpu_z_next_deformable.py "$@" -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_new.txt
echo "== timing"
timeout 120 python tools/next_bench.py > $OUT/${TAG}_next.json 2> $OUT/${TAG}_next.err; echo "next_bench exit $?"; tail -c 600 $OUT/${TAG}_next.err
cat $OUT/${TAG}_next.json
echo "== ncu --set full"
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"sgemm_kernel|voxel_lbs|colsum|split_reduce" -s 60 -c 40 -o $OUT/${TAG}_next_full \
    python tools/next_bench.py 50000 8 > $OUT/${TAG}_ncu_next.log 2>&1
ls -la $OUT/${TAG}_next_full.ncu-rep 2>&1
echo done
