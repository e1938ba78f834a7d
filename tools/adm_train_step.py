"""ImageNet-64 ADM U-Net (configs/imagenet64/T10.yaml) forward + backward on the B200 plan vs the same graph in eager PyTorch
(fp16 torso under autograd, as the reference trains it) on the same GPU.  Usage: python tools/adm_train_step.py [B ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import EDM_IN64_CFG, adm_oracle_kwargs, build_edm  # noqa: E402
from oracle import nets  # noqa: E402  (the eager baseline: test / measurement infrastructure only)

Bs = [int(a) for a in sys.argv[1:]] or [16, 64]
unet, sampler, sd = build_edm(EDM_IN64_CFG, T=10, fp16=True)
unet.train()
dev = "cuda"
torch.backends.cudnn.benchmark = True
for B in Bs:
    g = torch.Generator().manual_seed(B)
    x = torch.randn(B, 3, 64, 64, generator=g).to(dev)
    t = (torch.randn(B, generator=g) * 1.2 - 0.4).to(dev)
    y = torch.randint(0, 1000, (B,), generator=g).to(dev)
    coef = torch.randn(B, 3, 64, 64, generator=g).to(dev)

    def step():
        for p in unet.parameters():
            p.grad = None
        out = unet(x, t, y)
        (out * coef).sum().backward()
        return out

    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = int(os.environ.get("ADM_STEP_ITERS", "5"))
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    gn = torch.sqrt(sum((p.grad.float() ** 2).sum() for p in unet.parameters() if p.grad is not None)).item()
    print(f"B={B}: B200 plan forward+backward {ms:.1f} ms ({B / ms * 1e3:.0f} img/s), grad norm {gn:.3e}, "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB", flush=True)
    # eager baseline: fp16 torso autograd (the reference's training numerics) with the same weights
    rsd = {}
    torso = ("input_blocks", "middle_block", "output_blocks")
    for k, v in sd.items():
        if k == "log_betas":
            continue
        v = (v[..., None] if v.dim() == 3 else v).to(dev)
        is_conv = v.dim() == 4 or (k.endswith(".bias") and sd[k[:-5] + ".weight"].dim() >= 3)
        if k.split(".")[0] in torso and is_conv:
            v = v.half()
        rsd[k] = v.clone().requires_grad_(True)

    def eager():
        for v in rsd.values():
            v.grad = None
        with torch.device(dev):  # the oracle creates its small constants on the default device
            o = nets.adm_unet_forward(rsd, x, t, y, fp16_torso=True, **adm_oracle_kwargs(EDM_IN64_CFG))
        (o * coef).sum().backward()
        return o

    try:
        if os.environ.get("ADM_STEP_NO_EAGER"):
            raise torch.OutOfMemoryError("skipped")
        for _ in range(2):
            ref = eager()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            eager()
        e1.record()
        torch.cuda.synchronize()
        ems = e0.elapsed_time(e1) / n
        rel = ((out.detach() - ref.detach().float()).norm() / ref.detach().float().norm()).item()
        print(f"B={B}: eager PyTorch fp16-torso autograd {ems:.1f} ms -> {ems / ms:.2f}x; F rel-L2 vs eager {rel:.2e}", flush=True)
    except torch.OutOfMemoryError as e:  # noqa: PERF203
        print(f"B={B}: eager baseline out of memory ({e})")
    del rsd
    torch.cuda.empty_cache()
