N=${1:-8}
for algo in default NVLS NVLSTree Ring Tree; do
  if [ $algo = default ]; then unset NCCL_ALGO; else export NCCL_ALGO=$algo; fi
  echo "== NCCL_ALGO=$algo"; timeout 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/ar_probe.py 2>/dev/null | grep '^{"world"' | tee -a gpurun_out/r01o_ar${N}.json
done
unset NCCL_ALGO
for mode in none finish; do
  echo "== allreduce mode $mode"
  EMD_BENCH_ALLREDUCE=$mode python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r01o_bench_${N}gpu_$mode.json 2> gpurun_out/r01o_bench_${N}gpu_$mode.err; tail -c 300 gpurun_out/r01o_bench_${N}gpu_$mode.err; python tools/show_bench.py gpurun_out/r01o_bench_${N}gpu_$mode.json
done
nproc; free -g | head -2
