"""Determinism / A-B check of the inference rollout at a small batch: eager twice, graph replay, first_tc on/off."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
torch.set_grad_enabled(False)
from common import build_ddpm
from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.graph import GraphedRollout

def rel(a, b): return ((a.float() - b.float()).norm() / b.float().norm()).item()
for opt in (1, 0):
    L.lib().dxmi_set_option(b"first_tc", opt)
    net, sampler, value, sd, vsd = build_ddpm(10, device="cuda")
    B = 4
    torch.manual_seed(0)
    noise = torch.randn(11, B, 3, 32, 32, device="cuda")
    d1 = sampler.sample(B, device="cuda", noise=noise); e1 = value(d1["sample"], 10)
    d2 = sampler.sample(B, device="cuda", noise=noise); e2 = value(d2["sample"], 10)
    print("first_tc", opt, "eager twice equal:", torch.equal(d1["sample"], d2["sample"]), torch.equal(e1, e2))
    eps1 = net(noise[0], torch.full((B,), 3, device="cuda")); eps2 = net(noise[0], torch.full((B,), 3, device="cuda"))
    print("  single forward twice equal:", torch.equal(eps1, eps2))
    gr = GraphedRollout(sampler, B, "cuda", value=value)
    dg, eg = gr(noise); torch.cuda.synchronize()
    print("  graph equals eager:", torch.equal(dg["sample"], d1["sample"]), torch.equal(eg, e1), "rel", rel(dg["sample"], d1["sample"]))
    if opt == 1: keep = d1["sample"].clone()
    else: print("  tc vs fma rel-L2:", rel(keep, d1["sample"]))
