#!/bin/bash
# Short GPU visit for the SURVEY 8f-4 components: their parity tests first (plus the node tests whose host code changed),
# then the rest of the GPU suite, smoke(), the timing tool, a short bench line and an ncu capture of the new kernels.
# Usage (under gpurun): bash tools/gpu_s.sh <tag>
set -u
TAG=${1:-s}
OUT=gpurun_out
mkdir -p $OUT
echo "== new parity tests"
timeout 240 python -m pytest tests/test_gpu_z_next_voxel_lbs.py tests/test_gpu_z_next_deformable.py tests/test_gpu_emd_smpl.py \
    tests/test_gpu_emd_rigid.py -q -m gpu 2>&1 | tail -25 | tee $OUT/${TAG}_pytest_new.txt
echo "== timing of the new kernels"
timeout 120 python tools/next_bench.py > $OUT/${TAG}_next.json 2> $OUT/${TAG}_next.err; echo "next_bench exit $?"; tail -c 600 $OUT/${TAG}_next.err
cat $OUT/${TAG}_next.json
echo "== rest of the suite (+ smoke in parallel)"
( timeout 150 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.txt ) &
timeout 330 python -m pytest tests -m gpu -q --deselect tests/test_gpu_z_next_voxel_lbs.py --deselect tests/test_gpu_z_next_deformable.py \
    --deselect tests/test_gpu_emd_smpl.py --deselect tests/test_gpu_emd_rigid.py 2>&1 | tail -12 | tee $OUT/${TAG}_pytest_rest.txt
wait
tail -3 $OUT/${TAG}_smoke.txt
echo "== bench"
timeout 150 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 400 $OUT/${TAG}_bench.err
python tools/show_bench.py $OUT/${TAG}_bench.json 2>&1 | head -12
echo "== ncu --set full of the new kernels"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:"sgemm_kernel|voxel_lbs" -c 8 -o $OUT/${TAG}_next_full \
    python tools/next_bench.py 20000 4 > $OUT/${TAG}_ncu_next.log 2>&1
ls -la $OUT/${TAG}_next_full.ncu-rep 2>&1
echo done
