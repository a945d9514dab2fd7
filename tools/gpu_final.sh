#!/bin/bash
# Round-end check: the whole GPU suite (new files first), smoke(), one bench line.
set -u
TAG=${1:-w}
OUT=gpurun_out
mkdir -p $OUT
( timeout 150 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.txt 2>&1; echo "smoke exit $?" >> $OUT/${TAG}_smoke.txt ) &
timeout 400 python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee $OUT/${TAG}_pytest.txt
wait
tail -2 $OUT/${TAG}_smoke.txt
timeout 200 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
echo "bench exit $?"; tail -c 300 $OUT/${TAG}_bench.err
python tools/show_bench.py $OUT/${TAG}_bench.json 2>&1 | head -8
echo done
