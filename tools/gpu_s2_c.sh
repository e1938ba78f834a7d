#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_edm_gpu.py tests/test_adm_train_gpu.py tests/test_kernels_gpu.py -x -q 2>&1 | tail -4
python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s2_c_in64.json 2> gpurun_out/s2_c_in64.err
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/s2_c_cifar.json 2> gpurun_out/s2_c_cifar.err
python - <<'PY'
import json
for k in ("in64", "cifar"):
    try:
        d=json.loads(open(f"gpurun_out/s2_c_{k}.json").read().strip().splitlines()[-1])
        print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
    except Exception as e:
        print(k, "failed", e, open(f"gpurun_out/s2_c_{k}.err").read()[-800:])
PY
python tools/gemm_table.py --workload in64 > gpurun_out/s2_gemm_table_in64.txt 2>&1; head -30 gpurun_out/s2_gemm_table_in64.txt
