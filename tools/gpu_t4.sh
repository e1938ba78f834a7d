#!/bin/bash
mkdir -p gpurun_out

timeout -s KILL 900 python bench.py > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"; tail -3 gpurun_out/r2_bench1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench1.json'))
def show(x, tag):
    r=x['roofline']
    print(tag, 'value %.0f e2e %.0f ms %.2f | tensor frac %.3f share %.2f whole %.3f' % (x['value'], x['e2e']['value'], x['ms_per_step'], r['frac'], r['share_of_step'], r['whole_step_frac']))
    for h in r['hbm_kernels']: print('    hbm', h['kernel'][:40], 'GB/s %.0f frac %.2f share %.3f' % (h['achieved'], h['frac'], h['share_of_step']))
    if x.get('eager_cuda_baseline'): print('    eager', x['eager_cuda_baseline']['value'], x['eager_cuda_baseline']['mode'])
show(d,'main')
for s in d['secondary']: show(s, s['config']['workload'][:30])
print('cpu', d['cpu_baseline'])
PY
