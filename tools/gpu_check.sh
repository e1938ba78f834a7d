#!/bin/bash
# Runs the GPU parity suites under hard timeouts; logs land in gpurun_out/. Usage: tools/gpu_check.sh [test files...]
mkdir -p gpurun_out
FILES=${@:-tests/test_kernels_gpu.py tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_fp32_gpu.py}
for f in $FILES; do
  n=$(basename $f .py)
  timeout -s KILL 900 python -m pytest $f -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/$n.log 2>&1
  echo "$n rc=$?" | tee -a gpurun_out/$n.log
  tail -40 gpurun_out/$n.log
done
