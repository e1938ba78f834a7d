#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests/test_kernels_gpu.py tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_train_gpu.py tests/test_fp32_gpu.py -x -q 2>&1 | tail -8
for v in "a" "b pair_min=64" "c pair_min=32"; do
  set -- $v
  python bench.py --no-secondary --no-eager-baseline --no-cpu-baseline --steps 10 --warmup 3 ${2:+--opt $2} > gpurun_out/s2_b_$1.json 2> gpurun_out/s2_b_$1.err
done
python - <<'PY'
import json
for k in "abc":
    try:
        d=json.loads(open(f"gpurun_out/s2_b_{k}.json").read().strip().splitlines()[-1])
        print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
    except Exception as e:
        print(k, "failed", e, open(f"gpurun_out/s2_b_{k}.err").read()[-800:])
PY
bash tools/gpu_launch_list.sh | head -24
