#!/bin/bash
# ncu launch list of one profiled rollout -> gpurun_out/launches.csv + aggregated shares
mkdir -p gpurun_out
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_rollout.py --batch ${B:-256} --T ${T:-4} --rollouts 1 --warmup 1 --workload ${WL:-cifar} > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"
python tools/launch_shares.py gpurun_out/launches.csv | tee gpurun_out/launch_shares.txt | head -40
