"""The "real bar" of SURVEY 8d: the reference's arithmetic in eager PyTorch on the same B200 (the oracle's functional networks
moved to CUDA; TF32 and autocast-bf16), for the work of one CIFAR T=4 rollout at batch 256: 4 DDPM U-Net forwards + 1 value-net
forward (the transitions are a few elementwise launches).  Baseline only - nothing here is part of the product path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import json  # noqa: E402

import torch  # noqa: E402

from oracle import nets, synth  # noqa: E402

B, T = (int(sys.argv[1]) if len(sys.argv) > 1 else 256), 4
shapes = json.load(open(os.path.join(ROOT, "tests", "golden", "ddpm_shapes.json")))
sd = {k: v.cuda() for k, v in synth.synth_state_dict({k: tuple(v) for k, v in shapes["net"].items()}).items()}
vsd = {k: v.cuda() for k, v in synth.synth_state_dict({k: tuple(v) for k, v in shapes["value"].items()}, seed=1).items()}
torch.set_default_device("cuda")
x = torch.randn(B, 3, 32, 32)
t = torch.full((B,), 170.3)
torch.backends.cudnn.allow_tf32 = True
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.benchmark = True


def work(autocast):
    with torch.no_grad():
        h = x
        for _ in range(T):
            if autocast:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    eps = nets.ddpm_unet_forward(sd, h, t)
            else:
                eps = nets.ddpm_unet_forward(sd, h, t)
            h = 0.9 * h + 0.1 * eps.float()
        if autocast:
            with torch.autocast("cuda", dtype=torch.bfloat16):
                return nets.value_forward(vsd, h)
        return nets.value_forward(vsd, h)


for name, ac in (("fp32/TF32", False), ("autocast-bf16", True)):
    for _ in range(3):
        work(ac)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        work(ac)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"torch eager {name:14s} CIFAR T={T} rollout work, B={B}: {ms:8.2f} ms  {B / ms * 1e3:8.0f} img/s")
