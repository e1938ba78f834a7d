#!/bin/bash
# round-2 starting point: per-op timings + per-shape GEMM tables (CIFAR, IN64) + default bench
mkdir -p gpurun_out
timeout -s KILL 300 python tools/op_profile.py cifar 256 > gpurun_out/r2_opprof_cifar.log 2>&1; echo "opprof cifar rc=$?"; tail -1 gpurun_out/r2_opprof_cifar.log
python tools/op_times.py gpurun_out/ops_cifar.csv > gpurun_out/r2_op_times_cifar.txt; cat gpurun_out/r2_op_times_cifar.txt
timeout -s KILL 300 python tools/op_profile.py in64 64 > gpurun_out/r2_opprof_in64.log 2>&1; echo "opprof in64 rc=$?"; tail -1 gpurun_out/r2_opprof_in64.log
python tools/op_times.py gpurun_out/ops_in64.csv > gpurun_out/r2_op_times_in64.txt; cat gpurun_out/r2_op_times_in64.txt
timeout -s KILL 300 python tools/gemm_table.py --workload cifar > gpurun_out/r2_gemm_table_cifar.txt 2>&1; head -40 gpurun_out/r2_gemm_table_cifar.txt
timeout -s KILL 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; echo "bench rc=$?"; cat gpurun_out/r2_bench0.json
