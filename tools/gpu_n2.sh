#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_n2.json 2> gpurun_out/r2_n2.err; echo "n2 rc=$?"; tail -5 gpurun_out/r2_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_n2.json'))
print('main', d['value'], d['e2e']['value'], d['shard_check'], d['config']['collective'])
for s in d['secondary']: print('sec', s['value'], s['e2e']['value'], s['shard_check'])
print('c4', d['training_config']['value'], d['training_config']['ms_per_step'], d['training_config']['phases_ms'])
PY
DXMI_GRAPH_GATHER=1 timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --no-secondary > gpurun_out/r2_n2g.json 2> gpurun_out/r2_n2g.err; echo "n2 graph-gather rc=$?"; tail -3 gpurun_out/r2_n2g.err
python -c "
import json
d=json.load(open('gpurun_out/r2_n2g.json')); print('graph gather', d['value'], d['e2e']['value'], d['shard_check'])"
timeout -s KILL 600 python -m pytest tests/test_robustness_gpu.py tests/test_train_gpu.py -q -m gpu -k "non_current or ddp or DDP" 2>&1 | grep -E "Error|error|passed|failed" | tail -12
