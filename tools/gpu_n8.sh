#!/bin/bash
# 8-GPU visit: the bench at N=8 with the deferred SH-gradient exchange on / off, then configs[3] (6 M Gaussians, 5 cameras).
set -u
TAG=${1:-n8}
OUT=gpurun_out
mkdir -p $OUT
run() {  # name, env, extra args
  local name=$1; shift; local envs=$1; shift
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline "$@" > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "== $name exit $?"; tail -c 300 $OUT/${TAG}_${name}.err | tail -3; python tools/show_bench.py $OUT/${TAG}_${name}.json 2>&1 | head -3
}
run defer1 "EMD_BENCH_DEFER=1"
run defer0 "EMD_BENCH_DEFER=0"
run none "EMD_BENCH_ALLREDUCE=none"
run cfg3 "EMD_BENCH_DEFER=1" --n-bg 5800000 --cameras 5 --steps 10
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_n1.json 2>/dev/null; python tools/show_bench.py $OUT/${TAG}_n1.json | head -2
echo done
