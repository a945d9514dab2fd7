#!/bin/bash
# One GPU-box visit: parity tests, bench line, ncu launch list, ncu --set full of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip_full]
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
echo "== pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
echo "== bench"
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
tail -c 600 $OUT/${TAG}_bench.err
python - <<PY
import json
l=json.loads(open("$OUT/${TAG}_bench.json").read().strip().splitlines()[-1])
print("value",l["value"],"e2e",l["e2e"]["value"],"ms",l["ms_per_step"],"fwd_ms",l["fwd_ms_per_frame"],"clk",l["clocks"])
for k,v in l["roofline"]["per_kernel"].items(): print(f"  {k:16s} {v['ms_per_step']:8.4f} ms  x{v['launches_per_step']:.0f}  {v.get('achieved_gbs','')}")
print("cpu",l.get("cpu_baseline",{}).get("value"))
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
wc -l $OUT/${TAG}_launches.csv
if [ "${2:-}" != "skip_full" ]; then
echo "== ncu --set full (raster_bwd, raster_fwd_seg, onesweep, projection_bwd, gather, activate_bwd)"
timeout 1500 ncu --set full --clock-control none --import-source on \
    -k regex:'raster_bwd_kernel|raster_fwd_seg_kernel|rs_onesweep|projection_bwd|projection_fwd|raster_gather|activate_bwd|activate_fwd|isect_emit' \
    -s 60 -c 20 -o $OUT/${TAG}_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1
ls -la $OUT/${TAG}_full.ncu-rep
fi
echo '== s3g'
timeout 300 python tools/s3g_bench.py 1000000 > $OUT/${TAG}_s3g.json 2>&1; tail -c 1500 $OUT/${TAG}_s3g.json
echo done
