#!/bin/bash
# Session re-entry validation: whole GPU suite, smoke(), default bench line.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -5 > gpurun_out/s2_tests.txt; cat gpurun_out/s2_tests.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2_smoke.log 2>&1; echo "smoke rc=$?"
timeout -s KILL 900 python bench.py > gpurun_out/s2_bench.json 2> gpurun_out/s2_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
for s in d.get("secondary", []): print(s.get("config"), s.get("value"), s.get("ms_per_step"))
print(d.get("training_config"))
PY
