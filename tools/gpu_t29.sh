#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -3
timeout -s KILL 300 python tools/gemm_table.py --workload cifar > gpurun_out/t32_gemm_table_cifar.txt 2>&1; head -3 gpurun_out/t32_gemm_table_cifar.txt; grep -E "^ +262144 +8 |^ +1 +4992" gpurun_out/t32_gemm_table_cifar.txt
python bench.py --no-secondary --no-eager-baseline --steps 10 --warmup 3 > gpurun_out/t32_bench_a.json 2> gpurun_out/t32_bench_a.err
python - <<PY
import json
d=json.loads(open("gpurun_out/t32_bench_a.json").read().strip().splitlines()[-1])
print("a", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["whole_step_frac"], d["e2e"]["value"])
PY
