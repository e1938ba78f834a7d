#!/bin/bash
mkdir -p gpurun_out
python tools/bench_n128.py quick > gpurun_out/t11_n128.txt 2>&1
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_ddpm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu > gpurun_out/t11_tests.txt 2>&1
tail -3 gpurun_out/t11_tests.txt
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 > gpurun_out/t11_bench_a.json 2> gpurun_out/t11_bench_a.err
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 --opt s3_stages_max=3 > gpurun_out/t11_bench_b.json 2> gpurun_out/t11_bench_b.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --steps 5 --warmup 3 > gpurun_out/t11_bench_c.json 2> gpurun_out/t11_bench_c.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --steps 5 --warmup 3 --opt s3_stages_max=3 > gpurun_out/t11_bench_d.json 2> gpurun_out/t11_bench_d.err
for f in a b c d; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t11_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value"))
except Exception as e: print("$f", "ERR", e)
PY
done
tail -30 gpurun_out/t11_n128.txt
