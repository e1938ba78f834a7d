"""Aggregate the DXMI_TIME_OPS CSV (CUDA events around every plan op, warm caches, in pipeline order) by op kind."""
import collections
import re
import sys

agg = collections.defaultdict(lambda: [0, 0.0])
tot = 0.0
for line in open(sys.argv[1]):
    label, us = line.rsplit(",", 1)
    us = float(us)
    if " gemm " in label:
        m = re.search(r"M=(\d+) N=(\d+)", label)
        kind = f"GEMM M={m.group(1)} N={m.group(2)}"
    elif label.startswith("GN "):
        kind = "GroupNorm (finalize + apply / stats + apply)"
    elif label.startswith("ATTN "):
        kind = "fused attention"
    else:
        kind = "other (embed MLP, first conv, pool / upsample, attn_small)"
    agg[kind][0] += 1
    agg[kind][1] += us
    tot += us
for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:10.1f} us {100 * us / tot:5.1f}% n={n:5d} avg={us / n:8.1f}  {k}")
print(f"total {tot:.1f} us")
