#!/bin/bash
# ncu --set full captures at the end of round 2: (1) ImageNet-64 attention + qkv / proj_out projections + an up2 phase convolution,
# (2) CIFAR: pair GEMM N=128, the up2 GEMM, GroupNorm apply, fused attention block, last conv.  Text summaries only.
mkdir -p gpurun_out /tmp/ncu
timeout -s KILL 500 ncu --set full --clock-control none --import-source on -k regex:"attn_fwd_kernel|attn_small_k" -c 4 -f -o /tmp/ncu/s2_in64_attn \
    python tools/profile_rollout.py --workload in64 --batch 64 --T 1 --rollouts 1 --warmup 0 > gpurun_out/s2_ncu_in64.log 2>&1; echo "ncu in64 rc=$?"
python tools/ncu_summary.py /tmp/ncu/s2_in64_attn.ncu-rep > gpurun_out/s2_ncu_in64_attn_summary.txt 2>&1
timeout -s KILL 500 ncu --set full --clock-control none -k regex:"conv_gemm2p_kernel|gn_apply_ab_k|attnblk256_kernel|conv3x3_last_k|attn_small_k" -s 40 -c 60 -f -o /tmp/ncu/s2_cifar \
    python tools/profile_rollout.py --batch 256 --T 1 --rollouts 1 --warmup 0 > gpurun_out/s2_ncu_cifar.log 2>&1; echo "ncu cifar rc=$?"
python tools/ncu_summary.py /tmp/ncu/s2_cifar.ncu-rep > gpurun_out/s2_ncu_cifar_summary.txt 2>&1
grep -c "^---" gpurun_out/s2_ncu_in64_attn_summary.txt gpurun_out/s2_ncu_cifar_summary.txt
