#!/bin/bash
# `ncu --set full` over the hot kernels of a CIFAR T=4 B=256 rollout: (1) GEMM + GroupNorm kernels of the first U-Net level,
# (2) transition / attention / first conv.  Only the text summaries travel back (the .ncu-rep files are too large).
mkdir -p gpurun_out /tmp/ncu
timeout -s KILL 600 ncu --set full --clock-control none \
    -k regex:"conv_gemm2p_kernel|conv_gemm2_kernel|gn_apply_ab_k|gn_finalize_k|gn_stats_k|gn_apply_k" \
    -s ${1:-0} -c ${2:-30} -f -o /tmp/ncu/prof_gemm_gn \
    python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 0 > gpurun_out/ncu_kernels1.log 2>&1
echo "ncu1 rc=$?"
python tools/ncu_summary.py /tmp/ncu/prof_gemm_gn.ncu-rep > gpurun_out/prof_gemm_gn_summary.txt 2>&1
timeout -s KILL 600 ncu --set full --clock-control none \
    -k regex:"var_step_k|attn256_kernel|conv3x3_first_k|value_head_k|upsample2x_k" -c 8 -f -o /tmp/ncu/prof_misc \
    python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 0 > gpurun_out/ncu_kernels2.log 2>&1
echo "ncu2 rc=$?"
python tools/ncu_summary.py /tmp/ncu/prof_misc.ncu-rep > gpurun_out/prof_misc_summary.txt 2>&1
ls -la /tmp/ncu; grep -c "^---" gpurun_out/prof_gemm_gn_summary.txt gpurun_out/prof_misc_summary.txt
