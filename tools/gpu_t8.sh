#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_fullsize_gpu.py tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_train_gpu.py tests/test_backward_gpu.py -q -m gpu 2>&1 | tail -5
