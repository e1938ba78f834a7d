#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"gn_" -s 110 -c 56 --csv --log-file gpurun_out/gn_list.csv python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 1 > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/gn_list.csv') if not l.startswith('=='))]
by=collections.OrderedDict()
for r in rows:
    k=(r['ID'], r['Kernel Name'][:28], r['Grid Size'])
    by.setdefault(k,{})[r['Metric Name']]=float(r['Metric Value'].replace(',',''))
for k,v in by.items():
    print(k[1], k[2], 'us=%.1f rd=%.1fMB wr=%.1fMB' % (v['gpu__time_duration.sum']/ (1000 if v['gpu__time_duration.sum']>1000 else 1), v['dram__bytes_read.sum']/1e6 if v['dram__bytes_read.sum']>1e4 else v['dram__bytes_read.sum'], v['dram__bytes_write.sum']/1e6 if v['dram__bytes_write.sum']>1e4 else v['dram__bytes_write.sum']))
PY
head -3 gpurun_out/gn_list.csv
