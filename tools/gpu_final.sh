#!/bin/bash
# Round-end style validation: all GPU suites, smoke(), default bench (both arms), ncu launch list of the bench command.
mkdir -p gpurun_out
bash tools/gpu_check.sh tests/test_kernels_gpu.py tests/test_ddpm_gpu.py tests/test_edm_gpu.py tests/test_fullsize_gpu.py tests/test_fp32_gpu.py tests/test_backward_gpu.py tests/test_train_gpu.py 2>&1 | grep -E "passed|failed|rc="
timeout -s KILL 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; cat gpurun_out/final_bench.json; tail -3 gpurun_out/final_bench.err
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/final_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import collections, csv, re
lines = [l for l in open("gpurun_out/final_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[nm][0] += 1; agg[nm][1] += v; tot += v
with open("gpurun_out/final_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph\n")
    f.write("# whole process (warm-up + timed + e2e + roofline passes); cold-cache serialised durations: compare SHARES\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us {100*t/tot:5.1f}% n={c:6d} avg={t/c:8.1f} us  {k}\n")
    f.write(f"total {tot:.1f} us over {len(rows)} launches\n")
print(open("gpurun_out/final_launch_shares.txt").read()[:1800])
PY
