import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from common import EDM_IN64_CFG, build_edm
B = int(os.environ.get("B", "64"))
unet, sampler, sd = build_edm(EDM_IN64_CFG, 2)
x = torch.randn(B, 3, 64, 64, device="cuda")
y = torch.randint(0, 1000, (B,), device="cuda")
out = unet(x, torch.full((B,), 100.0, device="cuda"), y)
torch.cuda.synchronize()
print("ok", float(out.abs().mean()))
