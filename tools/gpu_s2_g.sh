#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_edm_gpu.py tests/test_backward_gpu.py -x -q 2>&1 | tail -4
timeout 200 python tools/bench_1x1.py 2>&1 | grep "pair_opt=1 bn=0"
python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s2_g_a.json 2> gpurun_out/s2_g_a.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 5 --warmup 3 --opt lean_epi=0 > gpurun_out/s2_g_b.json 2> gpurun_out/s2_g_b.err
python - <<'PY'
import json
for k in "ab":
    d=json.loads(open(f"gpurun_out/s2_g_{k}.json").read().strip().splitlines()[-1])
    print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
PY
