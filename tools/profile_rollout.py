"""Minimal workload for ncu captures: one warm-up rollout + one profiled rollout of the bench workload
(CIFAR DDPM U-Net, T=4, batch 256, + value net).  Usage (under gpurun):
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
      python tools/profile_rollout.py [--batch 256] [--T 4] [--rollouts 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

torch.set_grad_enabled(False)

from common import build_ddpm  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--T", type=int, default=4)
ap.add_argument("--rollouts", type=int, default=1)
ap.add_argument("--warmup", type=int, default=1)
ap.add_argument("--workload", default="cifar")
args = ap.parse_args()

if args.workload == "cifar":
    net, sampler, value, sd, vsd = build_ddpm(args.T, device="cuda")
    noise = torch.randn(args.T + 1, args.batch, 3, 32, 32, device="cuda")

    def run():
        d = sampler.sample(args.batch, device="cuda", noise=noise)
        return value(d["sample"], args.T)
else:
    from common import EDM_IN64_CFG, build_edm

    unet, sampler, sd = build_edm(EDM_IN64_CFG, args.T)
    noise = torch.randn(args.T, args.batch, 3, 64, 64, device="cuda")
    x0 = torch.randn(args.batch, 3, 64, 64, device="cuda") * 80
    y = torch.randint(0, 1000, (args.batch,), device="cuda")

    def run():
        return sampler.sample(args.batch, device="cuda", i_class=y, x0=x0, noise=noise)["sample"]
for _ in range(args.warmup):
    run()
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("profiled_rollouts")
for _ in range(args.rollouts):
    e = run()
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("output mean", float(e.mean()))
