#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_adm_train_gpu.py -x -q 2>&1 | tail -4
python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 > gpurun_out/s2_d_a.json 2> gpurun_out/s2_d_a.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2_d_a.json").read().strip().splitlines()[-1])
print("a", d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
PY
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"conv3x3_last_k|conv3x3_first_k" -c 6 python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 0 2>&1 | grep -E "conv3x3|gpu__time" | head -12
