#!/bin/bash
# Short GPU visit: 8f-4 parity tests, timing, ncu --set full of the GEMM flavours (3 launches each), nothing else.
set -u
TAG=${1:-v}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity tests"
timeout 200 python -m pytest tests/test_gpu_z_next_voxel_lbs.py tests/test_gpu_z_next_deformable.py -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_new.txt
echo "== timing"
timeout 100 python tools/next_bench.py > $OUT/${TAG}_next.json 2> $OUT/${TAG}_next.err; echo "exit $?"; tail -c 300 $OUT/${TAG}_next.err
cat $OUT/${TAG}_next.json
echo "== ncu --set full (layer 1 forward, its dgrad and wgrad, the voxel kernels)"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"sgemm_kernel|voxel_lbs|embed_grad_kernel|deform_input" -s 120 -c 14 \
    -o $OUT/${TAG}_next_full python tools/next_bench.py 50000 8 > $OUT/${TAG}_ncu_next.log 2>&1
ls -la $OUT/${TAG}_next_full.ncu-rep 2>&1
echo done
