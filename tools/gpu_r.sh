#!/bin/bash
# Short GPU visit: bench line first (alone on the GPU), then parity tests and smoke() side by side, then the
# ncu launch list of the bench command.  Usage (under gpurun): bash tools/gpu_r.sh <tag> [pytest-args]
set -u
TAG=${1:-r}
OUT=gpurun_out
mkdir -p $OUT
echo "== bench"
timeout 420 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 800 $OUT/${TAG}_bench.err
python tools/show_bench.py $OUT/${TAG}_bench.json 2>&1 | head -40
echo "== pytest -m gpu  (+ smoke in parallel)"
( timeout 200 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.txt ) &
timeout 600 python -m pytest tests -m gpu -x -q ${2:-} 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt
wait
tail -3 $OUT/${TAG}_smoke.txt
echo "== ncu launch list"
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
wc -l $OUT/${TAG}_launches.csv
echo done
