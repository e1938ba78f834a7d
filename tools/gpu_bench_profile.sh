#!/bin/bash
# bench (N=1) + ncu launch list + one full ncu capture of the dominant kernel. Outputs in gpurun_out/.
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout -s KILL 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
echo "ref rc=$?"; cat gpurun_out/bench_ref.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 1 > gpurun_out/ncu_list.log 2>&1
echo "ncu list rc=$?"; wc -l gpurun_out/launches.csv
timeout -s KILL 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm -s 100 -c 8 -f -o gpurun_out/prof_conv \
    python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 1 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out
