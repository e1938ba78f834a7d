#!/bin/bash
mkdir -p gpurun_out
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 > gpurun_out/t27_bench_a.json 2> gpurun_out/t27_bench_a.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --steps 5 --warmup 3 > gpurun_out/t27_bench_c.json 2> gpurun_out/t27_bench_c.err
python bench.py --workload c4 --no-secondary --no-eager-baseline --steps 5 --warmup 3 > gpurun_out/t27_bench_d.json 2> gpurun_out/t27_bench_d.err
for f in a c d; do python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/t27_bench_$f.json").read().strip().splitlines()[-1])
    print("$f", d["value"], d["ms_per_step"], d.get("roofline",{}).get("frac"), d.get("roofline",{}).get("whole_step_frac"), d.get("e2e",{}).get("value"), d.get("phases_ms"))
except Exception as e: print("$f", "ERR", e)
PY
done
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -3
