"""Value-net training step on the B200 path: forward + backward through the public API (the drop-in TimeIndependentValue under
autograd), CUDA events, vs torch eager (fp32 and autocast-bf16) running the oracle's functional network on the same GPU.
Usage: python tools/bench_train.py [B] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import VALUE_CFG, load_synth_into  # noqa: E402
from oracle import nets  # noqa: E402  (baseline only)

from diffusion_by_maxentirl_b200 import _lib as L  # noqa: E402
from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2  # noqa: E402
from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 20
value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
vsd = load_synth_into(value, seed=1)
value.cuda()
x = torch.randn(B, 3, 32, 32, device="cuda")
FWD_GFLOP = 1.613 * B  # value net forward, 32x32


def timed(fn, n):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def ours(need_dx):
    xi = x.clone().requires_grad_(need_dx)

    def step():
        for p in value.parameters():
            p.grad = None
        out = value(xi, 0)
        out.sum().backward()

    return step


def torch_ref(autocast):
    sd = {k: v.cuda().requires_grad_(True) for k, v in vsd.items()}

    def step():
        for p in sd.values():
            p.grad = None
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                out = nets.value_forward(sd, x)
        else:
            out = nets.value_forward(sd, x)
        out.float().sum().backward()

    return step


n0 = L.lib().dxmi_launch_count()
ms = timed(ours(False), iters)
launches = (L.lib().dxmi_launch_count() - n0) / (iters + 3)
print(f"B200 path  fwd+bwd (param grads)      B={B}: {ms:7.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s  {B / ms * 1e3:9.0f} img/s  ({launches:.0f} launches)")
if len(sys.argv) > 3 and sys.argv[3] == "ours":
    sys.exit(0)
ms = timed(ours(True), iters)
print(f"B200 path  fwd+bwd (param + input)    B={B}: {ms:7.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s")
with torch.no_grad():
    ms = timed(lambda: value(x, 0), iters)
print(f"B200 path  forward only (inference)   B={B}: {ms:7.3f} ms  {FWD_GFLOP / ms:7.1f} TFLOP/s")
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
ms = timed(torch_ref(False), max(3, iters // 4))
print(f"torch eager fp32/TF32 fwd+bwd         B={B}: {ms:7.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s")
ms = timed(torch_ref(True), max(3, iters // 4))
print(f"torch eager autocast-bf16 fwd+bwd     B={B}: {ms:7.3f} ms  {3 * FWD_GFLOP / ms:7.1f} TFLOP/s")
