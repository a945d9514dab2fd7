#!/bin/bash
# Short GPU visit: parity tests of the 8f-4 components, then the timing tool once per sgemm register flavour.
set -u
TAG=${1:-u}
OUT=gpurun_out
mkdir -p $OUT
echo "== parity tests"
timeout 200 python -m pytest tests/test_gpu_z_next_voxel_lbs.py tests/test_gpu_z_next_deformable.py -q -m gpu 2>&1 | tail -15 | tee $OUT/${TAG}_pytest_new.txt
for MB in 2 1; do
  echo "== timing, EMD_SGEMM_MINB=$MB"
  EMD_SGEMM_MINB=$MB timeout 100 python tools/next_bench.py > $OUT/${TAG}_next_minb$MB.json 2> $OUT/${TAG}_next_minb$MB.err; echo "exit $?"; tail -c 300 $OUT/${TAG}_next_minb$MB.err
  python - $OUT/${TAG}_next_minb$MB.json <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))["deform_network"]
print({k: d[k] for k in ("step_ms", "dense_fwd_ms", "dense_bwd_ms", "other_ms", "fwd_tflops_fp32", "bwd_tflops_fp32")})
PY
done
echo "== same with the tests under MINB=1"
EMD_SGEMM_MINB=1 timeout 120 python -m pytest tests/test_gpu_z_next_deformable.py -q -m gpu 2>&1 | tail -3
echo done
