#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_robustness_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu -s 2>&1 | grep -v "^$" | tail -30
timeout -s KILL 900 python -m pytest tests/test_train_gpu.py -x -q -m gpu 2>&1 | tail -3
