#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/bench_n128.py m2 > gpurun_out/t12_n128.txt 2>&1
cat gpurun_out/t12_n128.txt
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ddpm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/t12_tests.txt 2>&1
tail -15 gpurun_out/t12_tests.txt
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 > gpurun_out/t12_bench_a.json 2> gpurun_out/t12_bench_a.err
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 --opt s3_m2=0 > gpurun_out/t12_bench_b.json 2> gpurun_out/t12_bench_b.err
for f in a b; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t12_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value"), d.get("shard_check"))
except Exception as e: print("$f", "ERR", e)
PY
done
