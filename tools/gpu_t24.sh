#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_train_gpu.py -x -q -m gpu > gpurun_out/t25_tests.txt 2>&1
tail -5 gpurun_out/t25_tests.txt
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 > gpurun_out/t25_bench_a.json 2> gpurun_out/t25_bench_a.err
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 --opt t_uniform=0 > gpurun_out/t25_bench_b.json 2> gpurun_out/t25_bench_b.err
for f in a b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t25_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d.get("e2e",{}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
