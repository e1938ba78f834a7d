#!/bin/bash
mkdir -p gpurun_out
for v in "a" "b pair_min=48" "c pair_min=16"; do
  set -- $v
  python bench.py --workload in64 --no-secondary --no-eager-baseline --no-cpu-baseline --steps 4 --warmup 3 ${2:+--opt $2} > gpurun_out/s2_i_$1.json 2> gpurun_out/s2_i_$1.err
done
python - <<'PY'
import json
for k in "abc":
    d=json.loads(open(f"gpurun_out/s2_i_{k}.json").read().strip().splitlines()[-1])
    print(k, d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
PY
