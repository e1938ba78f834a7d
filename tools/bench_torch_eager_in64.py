"""SURVEY 8d "real bar" for ImageNet-64: the oracle's functional ADM U-Net (fp16 torso like the reference) in eager PyTorch on the same
B200: 10 U-Net forwards at batch 64 (the T=10 rollout's work without the value net).  Baseline only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import EDM_IN64_CFG, adm_oracle_kwargs, build_edm  # noqa: E402
from oracle import nets  # noqa: E402

B, T = 64, 10
unet, _, _ = build_edm(EDM_IN64_CFG, T)  # synthetic weights, convert_to_fp16() applied: fp16 torso convs, fp32 norms / embeddings
sd = {k: (v[..., None] if v.dim() == 3 else v).detach() for k, v in unet.state_dict().items()}
akw = adm_oracle_kwargs(EDM_IN64_CFG)
torch.set_default_device("cuda")
torch.backends.cudnn.benchmark = True
x = torch.randn(B, 3, 64, 64)
t = torch.full((B,), 350.7)
y = torch.randint(0, 1000, (B,))


def work():
    with torch.no_grad():
        h = x
        for _ in range(T):
            F_ = nets.adm_unet_forward(sd, h, t, y, fp16_torso=True, **akw)
            h = 0.9 * h + 0.1 * F_.float()
        return h


work()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    work()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 2
print(f"torch eager fp16-torso ADM U-Net, ImageNet-64 T={T} rollout work, B={B}: {ms:8.1f} ms  {B / ms * 1e3:7.1f} img/s")
