#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/s2m_tests.txt; cat gpurun_out/s2m_tests.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2m_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/s2m_smoke.log
timeout -s KILL 900 python bench.py > gpurun_out/s2m_bench.json 2> gpurun_out/s2m_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open("gpurun_out/s2m_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"], d["cpu_baseline"]["value"], d["eager_cuda_baseline"]["value"])
for s in d.get("secondary", []): print(s.get("config", {}).get("workload", "")[:40], s.get("value"), s.get("ms_per_step"), s["e2e"]["value"], s["roofline"].get("frac"), s["roofline"].get("whole_step_frac"), s["eager_cuda_baseline"]["value"])
print((d.get("training_config") or {}).get("ms_per_step"), (d.get("training_config") or {}).get("phases_ms"))
print(d["roofline"]["hbm_kernels"][0]["frac"], d["roofline"]["hbm_kernels"][0]["share_of_step"], d["clocks"])
PY
