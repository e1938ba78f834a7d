"""DDPM U-Net training step on the B200 path: eps = net(x, t) forward + backward through the public API (drop-in Model in train()
mode under autograd), CUDA events, vs torch eager (TF32 and autocast-bf16) running the oracle's functional U-Net on the same GPU.
Usage: python tools/bench_train_unet.py [B] [iters] [ours]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import DDPM_CFG, load_synth_into  # noqa: E402
from oracle import nets  # noqa: E402  (baseline only)

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
net = Model(**dict(DDPM_CFG, dropout=0.0))
sd = load_synth_into(net)
net.cuda().train()
x = torch.randn(B, 3, 32, 32, device="cuda")
t = torch.full((B,), 170.3, device="cuda")
coef = torch.randn(B, 3, 32, 32, device="cuda")
FWD_GFLOP = 12.444 * B


def timed(fn, n):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ours():
    for p in net.parameters():
        p.grad = None
    (net(x, t) * coef).sum().backward()


n0 = L.lib().dxmi_launch_count()
ms = timed(ours, iters)
launches = (L.lib().dxmi_launch_count() - n0) / (iters + 2)
print(f"B200 path  U-Net fwd+bwd  B={B}: {ms:8.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s  {B / ms * 1e3:8.0f} img/s  ({launches:.0f} launches)")
net.eval()
with torch.no_grad():
    ms = timed(lambda: net(x, t), iters)
print(f"B200 path  U-Net forward (inference plan) B={B}: {ms:8.3f} ms  {FWD_GFLOP / ms:7.1f} TFLOP/s")
print(f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB (torch) ; arena {L.lib().dxmi_workspace_bytes(net._handle, B) / 2**30:.2f} GiB (inference plan)")
if len(sys.argv) > 3 and sys.argv[3] == "ours":
    sys.exit(0)
rsd = {k: v.cuda().requires_grad_(True) for k, v in sd.items()}
torch.set_default_device("cuda")  # the oracle builds its frequency table on the default device


def torch_ref(autocast):
    def step():
        for p in rsd.values():
            p.grad = None
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = nets.ddpm_unet_forward(rsd, x, t)
        else:
            out = nets.ddpm_unet_forward(rsd, x, t)
        (out.float() * coef).sum().backward()

    return step


torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
ms = timed(torch_ref(False), max(2, iters // 3))
print(f"torch eager fp32/TF32 U-Net fwd+bwd     B={B}: {ms:8.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s")
ms = timed(torch_ref(True), max(2, iters // 3))
print(f"torch eager autocast-bf16 U-Net fwd+bwd B={B}: {ms:8.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s")
