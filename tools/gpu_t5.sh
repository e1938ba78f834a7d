#!/bin/bash
mkdir -p gpurun_out
for o in wave_bn=0 wave_bn=1; do
timeout -s KILL 600 python bench.py --steps 10 --no-cpu-baseline --no-eager-baseline --no-secondary --opt $o > gpurun_out/r2_b.json 2> gpurun_out/r2_b.err; echo "$o rc=$?"; tail -2 gpurun_out/r2_b.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_b.json'))
for x in [d]+d['secondary']:
    r=x['roofline']
    print('value %.0f e2e %.0f ms %.2f tensor %.3f whole %.3f' % (x['value'], x['e2e']['value'], x['ms_per_step'], r['frac'], r['whole_step_frac']), [(round(h['achieved']), round(h['share_of_step'],3)) for h in r['hbm_kernels']])
PY
done
timeout -s KILL 300 python tools/gemm_table.py --workload cifar 2>&1 | grep -E "^ +65536 +256" 
timeout -s KILL 900 python -m pytest tests/test_fullsize_gpu.py tests/test_ddpm_gpu.py -q -m gpu -x 2>&1 | tail -2
