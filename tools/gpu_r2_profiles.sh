#!/bin/bash
# Round-2 evidence: default bench (both arms), ncu launch list of the bench command, ncu --set full summaries of the hot kernels.
mkdir -p gpurun_out/ncu
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r02_bench.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph > gpurun_out/r02_ncu_list.log 2>&1; echo "ncu list rc=$?"
python - <<'PY'
import collections, csv, re
lines = [l for l in open("gpurun_out/r02_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[nm][0] += 1; agg[nm][1] += v; tot += v
with open("gpurun_out/r02_launch_shares_bench_cmd.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph\n")
    f.write("# whole process (warm-up + timed + e2e + roofline passes); cold-cache serialised durations: compare SHARES\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us {100*t/tot:5.1f}% n={c:6d} avg={t/c:8.1f} us  {k}\n")
    f.write(f"total {tot:.1f} us over {len(rows)} launches\n")
print(open("gpurun_out/r02_launch_shares_bench_cmd.txt").read()[:2500])
PY
cap() { # name regex skip count
  timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -f -o gpurun_out/ncu/r02_$1 \
      python tools/profile_rollout.py --batch 256 --T 4 --rollouts 1 --warmup 1 > gpurun_out/ncu/r02_$1.log 2>&1
  echo "$1 rc=$?"
  python tools/ncu_summary.py gpurun_out/ncu/r02_$1.ncu-rep > gpurun_out/ncu/r02_$1_summary.txt 2>&1
}
cap gn "gn_apply_ab_k|gn_finalize_k|gn_apply_fused_k" 60 8
cap attnblk "attnblk256_kernel" 5 1
cap step "var_step_k|value_head_k|conv3x3_first_k" 4 3
cap gemm2p "conv_gemm2p_kernel" 70 8
cap gemm2 "conv_gemm2_kernel" 40 6
rm -f gpurun_out/ncu/r02_gemm2.ncu-rep gpurun_out/ncu/r02_step.ncu-rep
ls -la gpurun_out/ncu | head -30
