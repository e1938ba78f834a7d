"""Warm, in-pipeline per-op timing of one U-Net forward (CUDA events around every plan op). Usage: op_profile.py cifar|in64 [B]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import EDM_IN64_CFG, build_ddpm, build_edm  # noqa: E402

wl = sys.argv[1] if len(sys.argv) > 1 else "cifar"
if wl == "cifar":
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    net, sampler, value, sd, vsd = build_ddpm(4)
    x = torch.randn(B, 3, 32, 32, device="cuda")
    run = lambda: net(x, torch.full((B,), 100.0, device="cuda"))  # noqa: E731
else:
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    net, sampler, sd = build_edm(EDM_IN64_CFG, 2)
    x = torch.randn(B, 3, 64, 64, device="cuda")
    y = torch.randint(0, 1000, (B,), device="cuda")
    run = lambda: net(x, torch.full((B,), 100.0, device="cuda"), y)  # noqa: E731
for _ in range(3):
    run()
torch.cuda.synchronize()
path = os.path.join(ROOT, "gpurun_out", f"ops_{wl}.csv")
if os.path.exists(path):
    os.remove(path)
os.environ["DXMI_TIME_OPS"] = path
run()
torch.cuda.synchronize()
del os.environ["DXMI_TIME_OPS"]
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    run()
e1.record()
torch.cuda.synchronize()
print(f"{wl} B={B}: forward {e0.elapsed_time(e1) / 5:.2f} ms (eager)")
