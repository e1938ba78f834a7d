#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2z_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph > gpurun_out/s2z_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import collections, csv, re
lines = [l for l in open("gpurun_out/s2z_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[nm][0] += 1; agg[nm][1] += v; tot += v
with open("gpurun_out/s2z_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph  (final library of round 2)\n")
    f.write("# whole process (warm-up + timed + e2e + roofline passes); cold-cache serialised durations: compare SHARES\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us {100*t/tot:5.1f}% n={c:6d} avg={t/c:8.1f} us  {k}\n")
    f.write(f"total {tot:.1f} us over {len(rows)} launches\n")
print(open("gpurun_out/s2z_launch_shares.txt").read()[:1800])
PY
