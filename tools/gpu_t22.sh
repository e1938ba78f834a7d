#!/bin/bash
mkdir -p gpurun_out
ADM_STEP_ITERS=1 ADM_STEP_NO_EAGER=1 timeout 800 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t22_adm_launches.csv python tools/adm_train_step.py 64 > gpurun_out/t22_ncu.log 2>&1
tail -3 gpurun_out/t22_ncu.log
python - <<'PY'
import csv, collections, re
rows=[]
with open('gpurun_out/t22_adm_launches.csv') as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
tot=collections.defaultdict(lambda:[0.0,0])
for row in r:
    try:
        v=float(row['Metric Value'].replace(',',''))
    except Exception: continue
    unit=row.get('Metric Unit','')
    if unit in ('nsecond','ns'): v/=1000.0
    elif unit in ('msecond','ms'): v*=1000.0
    name=re.sub(r'\(.*','',row['Kernel Name'])[:60]
    tot[name][0]+=v; tot[name][1]+=1
s=sum(v[0] for v in tot.values())
out=[f"# ncu --metrics gpu__time_duration.sum: tools/adm_train_step.py 64 (2 warm-up + 1 timed forward+backward of the ImageNet-64 ADM U-Net, B=64); serialised cold-cache durations - compare SHARES\n# total {s/1000:.1f} ms over {sum(v[1] for v in tot.values())} launches"]
for k,v in sorted(tot.items(), key=lambda kv:-kv[1][0])[:40]:
    out.append(f"{v[0]:12.1f} us {100*v[0]/s:5.1f}% n={v[1]:6d} avg={v[0]/v[1]:8.1f} us  {k}")
open('gpurun_out/t22_adm_train_launch_shares.txt','w').write('\n'.join(out)+'\n')
print('\n'.join(out[:32]))
PY
