#!/bin/bash
# ncu launch list (durations only) of the value-net training step; shares per kernel -> gpurun_out/train_launch_shares.txt
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/train_launches.csv \
    python tools/bench_train.py ${1:-256} 2 ours > gpurun_out/train_ncu.log 2>&1
echo "ncu rc=$?"
python - > gpurun_out/train_launch_shares.txt <<'PY'
import collections, csv, re
lines = [l for l in open("gpurun_out/train_launches.csv") if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines)]
# 5 identical steps (3 warm-up + 2 timed) after the one-off weight packing: keep the last fifth
own = [r for r in rows if not r["Kernel Name"].startswith("void at::") and "pack_conv" not in r["Kernel Name"]]
n = len(own) // 5
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in own[-n:]:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:50]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[(nm, r["Grid Size"])][0] += 1; agg[(nm, r["Grid Size"])][1] += v; tot += v
print("# one value-net training step (forward + backward, B=256), ncu durations (cold cache, serialised): compare SHARES")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}% n={c:3d} avg={t/c:7.1f} us  {k[0]} grid={k[1]}")
print(f"total {tot:.1f} us over {n} launches")
PY
head -30 gpurun_out/train_launch_shares.txt
