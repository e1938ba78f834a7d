#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_train_gpu.py tests/test_robustness_gpu.py -x -q 2>&1 | tail -4
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/s2_f_a.json 2> gpurun_out/s2_f_a.err
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 --opt first_tc=0 > gpurun_out/s2_f_b.json 2> gpurun_out/s2_f_b.err
python - <<'PY'
import json
for k in "ab":
    d=json.loads(open(f"gpurun_out/s2_f_{k}.json").read().strip().splitlines()[-1])
    print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
PY
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3x3_first" -c 4 python tools/profile_rollout.py --batch 256 --T 2 --rollouts 1 --warmup 0 2>&1 | grep -E "gpu__time|conv3x3_first.*\(" | head -8
