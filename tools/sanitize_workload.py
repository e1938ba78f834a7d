"""Smallest end-to-end workload for compute-sanitizer: DDPM T=2 rollout at B=2 (+ value net) and one reduced-width EDM step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import EDM_SMALL_CFG, build_ddpm, build_edm  # noqa: E402

torch.set_grad_enabled(False)
net, sampler, value, sd, vsd = build_ddpm(2)
d = sampler.sample(2, device="cuda")
e = value(d["sample"], 2)
torch.cuda.synchronize()
print("ddpm ok", float(d["sample"].abs().mean()), float(e.mean()))
if "--edm" in sys.argv:
    unet, es, _ = build_edm(EDM_SMALL_CFG, 2)
    d = es.sample(2, "cuda", i_class=torch.tensor([1, 2], device="cuda"))
    torch.cuda.synchronize()
    print("edm ok", float(d["sample"].abs().mean()))
