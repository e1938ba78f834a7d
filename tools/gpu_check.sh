#!/bin/bash
# Runs the GPU parity suites under hard timeouts; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/kernels.log 2>&1
echo "kernels rc=$?" | tee -a gpurun_out/kernels.log
tail -40 gpurun_out/kernels.log
timeout -s KILL 900 python -m pytest tests/test_ddpm_gpu.py -m gpu -q -s --no-header -p no:cacheprovider > gpurun_out/ddpm.log 2>&1
echo "ddpm rc=$?" | tee -a gpurun_out/ddpm.log
tail -60 gpurun_out/ddpm.log
