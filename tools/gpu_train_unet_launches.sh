#!/bin/bash
# ncu launch list (durations only) of the DDPM U-Net training step; shares per kernel -> gpurun_out/unet_train_launch_shares.txt
mkdir -p gpurun_out
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/unet_train_launches.csv \
    python tools/bench_train_unet.py ${1:-128} 1 ours > gpurun_out/unet_train_ncu.log 2>&1
echo "ncu rc=$?"
python - > gpurun_out/unet_train_launch_shares.txt <<'PY'
import collections, csv, re
lines = [l for l in open("gpurun_out/unet_train_launches.csv") if not l.startswith("==")]
rows = [r for r in csv.DictReader(lines)]
own = [r for r in rows if "pack_conv" not in r["Kernel Name"] and "cast_to" not in r["Kernel Name"]]
# bench_train_unet: 3 identical training steps (2 warm-up + 1 timed), then inference forwards: locate the steps by their first kernel
starts = [i for i, r in enumerate(own) if "timestep_embedding" in r["Kernel Name"]]
seg = own[starts[2]:starts[3]] if len(starts) > 3 else own[starts[-1]:]
agg = collections.defaultdict(lambda: [0, 0.0]); tot = 0.0
for r in seg:
    nm = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("dxmi::", "")[:48]
    v = float(r["Metric Value"].replace(",", "")) / 1e3
    agg[nm][0] += 1; agg[nm][1] += v; tot += v
print("# one DDPM U-Net training step (forward + backward, B=128), ncu durations (cold cache, serialised): compare SHARES")
for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:9.1f} us {100*t/tot:5.1f}% n={c:4d} avg={t/c:7.1f} us  {k}")
print(f"total {tot:.1f} us over {len(seg)} launches")
PY
head -32 gpurun_out/unet_train_launch_shares.txt
