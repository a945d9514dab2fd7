#!/bin/bash
# tuning experiment: segment length of the segment-parallel raster kernels
set -u
for SB in 4 8 16; do
  echo "=== EMD_SEG_BATCHES=$SB"
  EMD_SEG_BATCHES=$SB python -m emd_b200.build --force > /dev/null 2>&1
  timeout 600 python -m pytest tests/test_gpu_rasterization.py tests/test_gpu_sort.py -m gpu -x -q 2>&1 | tail -2
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/seg_$SB.json 2>/dev/null
  python - <<PY
import json
l=json.loads(open("gpurun_out/seg_$SB.json").read().strip().splitlines()[-1])
k=l["roofline"]["per_kernel"]
print("value",l["value"],"ms",l["ms_per_step"],"fwd_ms",l["fwd_ms_per_frame"], {n:k[n]["ms_per_step"] for n in ("raster_fwd","raster_bwd","sort_scatter","sort_hist","raster_gather")})
PY
done
