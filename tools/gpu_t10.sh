#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_edm_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout -s KILL 300 python tools/op_profile.py in64 64 > gpurun_out/r2_opprof_in64.log 2>&1; tail -1 gpurun_out/r2_opprof_in64.log
python tools/op_times.py gpurun_out/ops_in64.csv | grep -E "attention|total"
timeout -s KILL 300 python tools/gemm_table.py --workload in64 2>&1 | grep -E " 9 " | head
