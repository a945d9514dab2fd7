#!/bin/bash
# GPU visit for K1e: HexPlane parity tests first, then the whole GPU suite, the S3G timing and an ncu capture
# of the HexPlane / MLP kernels.  Usage (under gpurun): bash tools/gpu_hex.sh <tag>
set -u
TAG=${1:-hex}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest hexplane"
timeout 600 python -m pytest tests/test_gpu_hexplane.py tests/test_gpu_optim.py -q 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_hex.txt
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== s3g"
timeout 300 python tools/s3g_bench.py 1000000 > $OUT/${TAG}_s3g.json 2> $OUT/${TAG}_s3g.err; tail -c 2000 $OUT/${TAG}_s3g.json; tail -c 500 $OUT/${TAG}_s3g.err
echo "== adam"
timeout 200 python tools/adam_bench.py > $OUT/${TAG}_adam.json 2> $OUT/${TAG}_adam.err; tail -c 1200 $OUT/${TAG}_adam.json; tail -c 300 $OUT/${TAG}_adam.err
echo "== ncu full (hexplane fwd/bwd of one step) + launch list of the S3G step"
timeout 420 ncu --set full --clock-control none --import-source on -k regex:'hexplane_fwd|hexplane_bwd' \
    -s 6 -c 2 -o $OUT/${TAG}_hex_full python tools/s3g_bench.py 1000000 > $OUT/${TAG}_ncu_hex.log 2>&1
ls -la $OUT/${TAG}_hex_full.ncu-rep
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 160 --csv --log-file $OUT/${TAG}_s3g_launches.csv \
    python tools/s3g_bench.py 1000000 > $OUT/${TAG}_ncu_s3g.log 2>&1
wc -l $OUT/${TAG}_s3g_launches.csv
echo done
