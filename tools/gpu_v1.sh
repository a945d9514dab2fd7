#!/bin/bash
# Lean GPU visit: HexPlane/Adam parity first, whole GPU suite, S3G + Adam timings, one bench line.
set -u
TAG=${1:-v1}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest hexplane/optim"
timeout 400 python -m pytest tests/test_gpu_hexplane.py tests/test_gpu_optim.py -q 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_hex.txt
echo "== pytest -m gpu"
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== s3g"
timeout 200 python tools/s3g_bench.py 1000000 > $OUT/${TAG}_s3g.json 2> $OUT/${TAG}_s3g.err; tail -c 2500 $OUT/${TAG}_s3g.json; tail -c 500 $OUT/${TAG}_s3g.err
echo "== adam"
timeout 120 python tools/adam_bench.py > $OUT/${TAG}_adam.json 2> $OUT/${TAG}_adam.err; tail -c 1200 $OUT/${TAG}_adam.json; tail -c 300 $OUT/${TAG}_adam.err
echo "== bench"
timeout 300 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
python tools/show_bench.py $OUT/${TAG}_bench.json 2>&1 | head -40
echo done
