"""Per-shape table of the tcgen05 GEMM launches of one rollout (CUDA events around every launch).
Usage (under gpurun): python tools/gemm_table.py [--workload cifar|in64] [--batch B] [--T T]"""
import argparse
import collections
import csv
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

torch.set_grad_enabled(False)

from common import EDM_IN64_CFG, build_ddpm, build_edm  # noqa: E402
from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="cifar")
ap.add_argument("--batch", type=int, default=None)
ap.add_argument("--T", type=int, default=None)
args = ap.parse_args()
lib = L.lib()
if args.workload == "cifar":
    T, B = args.T or 4, args.batch or 256
    net, sampler, value, sd, vsd = build_ddpm(T)
    noise = torch.randn(T + 1, B, 3, 32, 32, device="cuda")
    run = lambda: value(sampler.sample(B, device="cuda", noise=noise)["sample"], T)  # noqa: E731
else:
    T, B = args.T or 2, args.batch or 64
    unet, sampler, sd = build_edm(EDM_IN64_CFG, T)
    noise = torch.randn(T, B, 3, 64, 64, device="cuda")
    x0 = torch.randn(B, 3, 64, 64, device="cuda") * 80
    y = torch.randint(0, 1000, (B,), device="cuda")
    run = lambda: sampler.sample(B, device="cuda", i_class=y, x0=x0, noise=noise)  # noqa: E731
for _ in range(2):
    run()
torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"gemm_launches_{args.workload}.csv")
os.makedirs(os.path.dirname(path), exist_ok=True)
lib.dxmi_set_timing_dump(path.encode())
lib.dxmi_set_option(b"time_gemms", 1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
run()
torch.cuda.synchronize()
ms, fl, nl = C.c_double(), C.c_double(), C.c_longlong()
lib.dxmi_gemm_timing(C.byref(ms), C.byref(fl), C.byref(nl))
lib.dxmi_set_option(b"time_gemms", 0)
lib.dxmi_set_timing_dump(None)
e0.record()
run()
e1.record()
torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
for r in csv.DictReader(open(path)):
    k = (int(r["M"]), int(r["N"]), int(r["K"]), int(r["batch"]), int(r["block_n"]), int(r["persistent"]))
    agg[k][0] += 1
    agg[k][1] += float(r["us"])
    agg[k][2] += float(r["gflop"])
print(f"# {args.workload} T={T} B={B}: {nl.value} GEMM launches, {ms.value:.2f} ms, {fl.value / 1e12:.2f} TFLOP, "
      f"{fl.value / ms.value / 1e9:.0f} TFLOP/s; whole rollout {e0.elapsed_time(e1):.2f} ms")
print(f"{'M':>8} {'N':>5} {'K':>5} {'bat':>4} {'bn':>4} {'v2':>2} {'n':>4} {'tot us':>9} {'avg us':>8} {'TFLOP/s':>8} {'share':>6}")
for k, (n, us, gf) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[0]:8d} {k[1]:5d} {k[2]:5d} {k[3]:4d} {k[4]:4d} {k[5]:2d} {n:4d} {us:9.1f} {us / n:8.1f} {gf / us * 1e3 / 1e3:8.1f} {100 * us / (ms.value * 1e3):5.1f}%")
