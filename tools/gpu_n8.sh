#!/bin/bash
# 8-GPU visit: the bench at N=8 under the environment settings given as arguments ("name:ENV=VAL,ENV2=VAL2" ...), N=1 last.
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for cfg in "$@"; do
  name=${cfg%%:*}; envs=$(echo "${cfg#*:}" | tr ',' ' ')
  np=8; case $name in n4*) np=4;; n2*) np=2;; esac
  env $envs timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $np --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "== $name exit $?"; python tools/show_bench.py $OUT/${TAG}_${name}.json 2>/dev/null | head -2
done
timeout 200 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_n1.json 2>/dev/null; python tools/show_bench.py $OUT/${TAG}_n1.json all 2>/dev/null | head -12
echo done
