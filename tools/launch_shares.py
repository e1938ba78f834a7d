"""Aggregate an ncu gpu__time_duration launch list (second half = the profiled rollout) by kernel and grid."""
import collections
import csv
import re
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rows = list(csv.DictReader(lines))
n = len(rows)
agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for row in rows[n // 2:]:
    nm = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:60]
    v = float(row["Metric Value"].replace(",", "")) / 1e3
    key = (nm, row["Grid Size"])
    agg[key][0] += 1
    agg[key][1] += v
    tot += v
print("# second (profiled) rollout; times are cold-cache, serialised ncu durations: compare SHARES")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100 * t / tot:5.1f}% n={c:4d} avg={t / c:8.1f} us  {k[0]} grid={k[1]}")
print(f"total {tot:.1f} us over {n - n // 2} launches")
