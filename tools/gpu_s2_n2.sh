#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/s2_n2.json 2> gpurun_out/s2_n2.err; echo "n2 rc=$?"; tail -3 gpurun_out/s2_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s2_n2.json').read().strip().splitlines()[-1])
print('main', d['value'], d['ms_per_step'], d['e2e']['value'], d['shard_check'])
for s in d['secondary']: print('sec', s['value'], s['e2e']['value'], s['shard_check'])
print('c4', d['training_config']['value'], d['training_config']['ms_per_step'], d['training_config'].get('phases_ms'))
PY
