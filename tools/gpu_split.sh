#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 python -m pytest tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py -x -q -m gpu 2>&1 | tail -5
for sp in 1 2 4; do
  timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --opt rollout_split=$sp > gpurun_out/r2_split$sp.json 2> gpurun_out/r2_split$sp.err; echo "split $sp rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2_split$sp.json'));print($sp, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" || tail -5 gpurun_out/r2_split$sp.err
done
for sp in 1 2; do
  timeout -s KILL 400 python bench.py --workload in64 --steps 5 --warmup 3 --no-cpu-baseline --opt rollout_split=$sp > gpurun_out/r2_in64_split$sp.json 2> gpurun_out/r2_in64_split$sp.err; echo "in64 split $sp rc=$?"
  python -c "import json;d=json.load(open('gpurun_out/r2_in64_split$sp.json'));print($sp, d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])" || tail -5 gpurun_out/r2_in64_split$sp.err
done
