#!/bin/bash
# Round-end style run: whole GPU suite, smoke(), both bench arms with the driver's command line, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout -s KILL 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -3 > gpurun_out/s2f_tests.txt; cat gpurun_out/s2f_tests.txt
timeout -s KILL 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s2f_smoke.log 2>&1; echo "smoke rc=$?"
timeout -s KILL 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/s2f_bench_ref.json 2> gpurun_out/s2f_bench_ref.err; echo "ref rc=$?"
timeout -s KILL 900 python bench.py > gpurun_out/s2f_bench.json 2> gpurun_out/s2f_bench.err; echo "bench rc=$?"
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2f_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph > gpurun_out/s2f_ncu.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import collections, csv, json, re
d=json.loads(open("gpurun_out/s2f_bench.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"].get("whole_step_frac"), d["e2e"]["value"])
for s in d.get("secondary", []): print(s.get("config", {}).get("workload", "")[:40], s.get("value"), s.get("ms_per_step"))
print((d.get("training_config") or {}).get("ms_per_step"), (d.get("training_config") or {}).get("phases_ms"))
lines = [l for l in open("gpurun_out/s2f_launches.csv") if not l.startswith("==")]
rows = list(csv.DictReader(lines))
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in rows:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:60]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[nm][0] += 1; agg[nm][1] += v; tot += v
with open("gpurun_out/s2f_launch_shares.txt", "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-secondary --no-graph\n")
    f.write("# whole process (warm-up + timed + e2e + roofline passes); cold-cache serialised durations: compare SHARES\n")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{t:10.1f} us {100*t/tot:5.1f}% n={c:6d} avg={t/c:8.1f} us  {k}\n")
    f.write(f"total {tot:.1f} us over {len(rows)} launches\n")
print(open("gpurun_out/s2f_launch_shares.txt").read()[:2200])
PY
