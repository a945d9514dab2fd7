#!/bin/bash
# One GPU visit: runs the numbered steps given as arguments (see the case table), everything into gpurun_out/<tag>_*.
set -u
TAG=$1; shift
OUT=gpurun_out
mkdir -p $OUT
for stepname in "$@"; do
  echo "== $stepname"
  case $stepname in
    tc_dense)  timeout 90 python tools/tc_dense_check.py > $OUT/${TAG}_tc_dense.txt 2>&1; echo "exit $?"; tail -30 $OUT/${TAG}_tc_dense.txt ;;
    tc_dense_big) timeout 120 python tools/tc_dense_check.py 50000 > $OUT/${TAG}_tc_dense_big.txt 2>&1; echo "exit $?"; tail -30 $OUT/${TAG}_tc_dense_big.txt ;;
    probe)     timeout 120 python tools/fp32_probe.py > $OUT/${TAG}_fp32_probe.json 2> $OUT/${TAG}_fp32_probe.err; echo "exit $?"; cat $OUT/${TAG}_fp32_probe.json; tail -3 $OUT/${TAG}_fp32_probe.err ;;
    counters)  timeout 300 python tools/raster_counters.py > $OUT/${TAG}_raster_counters.json 2> $OUT/${TAG}_raster_counters.err; echo "exit $?"; cat $OUT/${TAG}_raster_counters.json; tail -3 $OUT/${TAG}_raster_counters.err ;;
    sweep)     timeout 900 python tools/stage_bench.py --sweep > $OUT/${TAG}_raster_sweep.jsonl 2> $OUT/${TAG}_raster_sweep.err; echo "exit $?"; tail -c 600 $OUT/${TAG}_raster_sweep.err; wc -l $OUT/${TAG}_raster_sweep.jsonl ;;
    tests)     timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $OUT/${TAG}_pytest.txt ;;
    tests_new) timeout 900 python -m pytest ${EMD_TESTS:-tests/test_gpu_at_size.py} -m gpu -q -x 2>&1 | tail -30 | tee $OUT/${TAG}_pytest_new.txt ;;
    smoke)     timeout 200 python __graft_entry__.py smoke > $OUT/${TAG}_smoke.txt 2>&1; echo "exit $?"; tail -3 $OUT/${TAG}_smoke.txt ;;
    bench)     timeout 400 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "exit $?"; tail -c 400 $OUT/${TAG}_bench.err; python tools/show_bench.py $OUT/${TAG}_bench.json all 2>&1 | head -40 ;;
    bench_nocpu) timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "exit $?"; tail -c 400 $OUT/${TAG}_bench.err; python tools/show_bench.py $OUT/${TAG}_bench.json all 2>&1 | head -40 ;;
    bench_s3g) timeout 400 python bench.py --workload s3g --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench_s3g.json 2> $OUT/${TAG}_bench_s3g.err; echo "exit $?"; tail -c 600 $OUT/${TAG}_bench_s3g.err; head -c 1500 $OUT/${TAG}_bench_s3g.json ;;
    ncu_list)  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 400 --csv --log-file $OUT/${TAG}_ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1; echo "exit $?"; wc -l $OUT/${TAG}_ncu_launches.csv ;;
    ncu_bwd)   timeout 600 ncu --set full --metrics sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_lsu.sum,smsp__sass_thread_inst_executed_op_fadd_pred_on.sum,smsp__sass_thread_inst_executed_op_fmul_pred_on.sum,smsp__sass_thread_inst_executed_op_ffma_pred_on.sum --clock-control none --import-source on -k regex:"raster_bwd_kernel|raster_gather" -s 6 -c 2 -o $OUT/${TAG}_raster_bwd_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_full.log 2>&1; echo "exit $?"; ls -la $OUT/${TAG}_raster_bwd_full.ncu-rep ;;
    ncu_fwd)   timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_fwd_seg|raster_sort_pack" -s 12 -c 4 -o $OUT/${TAG}_raster_fwd_full python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_fwd.log 2>&1; echo "exit $?"; ls -la $OUT/${TAG}_raster_fwd_full.ncu-rep ;;
    *) echo "running custom: $stepname"; timeout 900 bash -c "$stepname" ;;
  esac
done
echo done
