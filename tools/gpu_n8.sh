N=${1:-8}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/ar_probe.py 2>/dev/null | grep '^{"world"' > gpurun_out/r01n_ar${N}.json; cat gpurun_out/r01n_ar${N}.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r01n_bench_${N}gpu.json 2> gpurun_out/r01n_bench_${N}gpu.err; tail -c 300 gpurun_out/r01n_bench_${N}gpu.err; python tools/show_bench.py gpurun_out/r01n_bench_${N}gpu.json
