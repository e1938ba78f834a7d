#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ddpm_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -8
timeout -s KILL 300 python tools/attnblk_timeline.py 256 | head -18
timeout -s KILL 300 python tools/op_profile.py cifar 256 > gpurun_out/r2_opprof_cifar.log 2>&1; tail -1 gpurun_out/r2_opprof_cifar.log
grep ATTNBLK gpurun_out/ops_cifar.csv | head -3
