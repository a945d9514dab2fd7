#!/bin/bash
# Quick GPU visit: parity tests + one bench line.  Usage: bash tools/gpu_quick.sh <tag> [pytest-args]
set -u
TAG=${1:-q}
OUT=gpurun_out
mkdir -p $OUT
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q ${2:-} 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value",l["value"],"e2e",l["e2e"]["value"],"ms",l["ms_per_step"],"fwd_ms",l["fwd_ms_per_frame"],"clk",l["clocks"])
for k,v in l["roofline"]["per_kernel"].items(): print(f"  {k:16s} {v['ms_per_step']:8.4f} ms  x{v['launches_per_step']:.0f}  {v.get('achieved_gbs','')}")
PY
