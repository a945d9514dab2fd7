#!/bin/bash
# 8-GPU NCCL probe: all-reduce of the bench's gradient sizes under different algorithm choices.
set -u
TAG=${1:-nccl}
OUT=gpurun_out
mkdir -p $OUT
for cfg in "default:" "nvls:NCCL_ALGO=NVLS" "ring:NCCL_ALGO=Ring" "tree:NCCL_ALGO=Tree" "ctas32:NCCL_MAX_CTAS=32" "ctas8:NCCL_MAX_CTAS=8"; do
  name=${cfg%%:*}; envs=${cfg#*:}
  env $envs NCCL_DEBUG=WARN timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 \
      tools/ar_probe.py 2>$OUT/${TAG}_${name}.err | tail -1 > $OUT/${TAG}_${name}.json
  echo "== $name"; cat $OUT/${TAG}_${name}.json | cut -c1-600
done
NCCL_DEBUG=INFO timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 tools/ar_probe.py 2>&1 | grep -iE "nvls|algo|channels" | head -12
echo done
