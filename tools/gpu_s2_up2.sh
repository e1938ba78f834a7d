#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "up2" 2>&1 | tail -15
timeout -s KILL 900 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py -x -q 2>&1 | tail -8
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/s2_up2_a.json 2> gpurun_out/s2_up2_a.err
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 --opt up2=0 > gpurun_out/s2_up2_b.json 2> gpurun_out/s2_up2_b.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/s2_up2_c.json 2> gpurun_out/s2_up2_c.err
python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 5 --warmup 3 --opt up2=0 > gpurun_out/s2_up2_d.json 2> gpurun_out/s2_up2_d.err
python - <<'PY'
import json
for k in "abcd":
    try:
        d=json.loads(open(f"gpurun_out/s2_up2_{k}.json").read().strip().splitlines()[-1])
        print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
    except Exception as e:
        print(k, "failed", e, open(f"gpurun_out/s2_up2_{k}.err").read()[-800:])
PY
