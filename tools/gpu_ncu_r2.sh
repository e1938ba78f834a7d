#!/bin/bash
# round 2: ncu --set full on the HBM-bound family + the new fused attention block + a few GEMMs; reports come back (small -c)
mkdir -p gpurun_out/ncu
cap() { # name regex skip count
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu/$1 \
      python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 1 > gpurun_out/ncu/$1.log 2>&1
  echo "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/ncu/$1.ncu-rep > gpurun_out/ncu/$1_summary.txt 2>&1
}
cap gn_apply "gn_apply_ab_k" 60 6
cap gn_finalize "gn_finalize_k" 60 2
cap attnblk "attnblk256_kernel" 5 1
cap step "var_step_k|value_head_k|conv3x3_first_k" 4 3
cap gemm2p "conv_gemm2p_kernel" 70 6
ls -la gpurun_out/ncu
