#!/bin/bash
timeout -s KILL 900 python -m pytest tests/test_train_gpu.py tests/test_backward_gpu.py -q -m gpu -x 2>&1 | tail -4
timeout -s KILL 300 python tools/bench_train_unet.py 128 10 2>&1 | grep "B200 path"
timeout -s KILL 300 python bench.py --workload c4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['phases_ms'])"
