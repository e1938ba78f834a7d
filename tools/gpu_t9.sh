#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python tools/op_profile.py in64 64 > gpurun_out/r2_opprof_in64.log 2>&1; tail -1 gpurun_out/r2_opprof_in64.log
python tools/op_times.py gpurun_out/ops_in64.csv
timeout -s KILL 300 python tools/gemm_table.py --workload in64 > gpurun_out/r2_gemm_table_in64.txt 2>&1; head -30 gpurun_out/r2_gemm_table_in64.txt
