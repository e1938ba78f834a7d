"""Diagnostic: per-sample / border-vs-interior error of the value-net input gradient against fp32 torch on the GPU."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from common import VALUE_CFG, load_synth_into  # noqa: E402
from oracle import nets  # noqa: E402

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2  # noqa: E402
from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue  # noqa: E402

value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
vsd = load_synth_into(value, seed=1)
value.cuda()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(5)
x = torch.randn(B, 3, 32, 32, generator=g)
coef = torch.randn(B, generator=g).cuda()
sd = {k: v.cuda().requires_grad_(True) for k, v in vsd.items()}
xr = x.cuda().requires_grad_(True)
out = nets.value_forward(sd, xr)
(out.view(-1) * coef).sum().backward()
xc = x.cuda().requires_grad_(True)
o2 = value(xc, 0)
(o2.view(-1) * coef).sum().backward()
d, r = xc.grad, xr.grad
print("out", o2.view(-1)[:4].tolist(), out.view(-1)[:4].tolist())
print("total rel", ((d - r).norm() / r.norm()).item())
for n in range(B):
    print(n, "rel", ((d[n] - r[n]).norm() / r[n].norm()).item(), "coef", coef[n].item())
e = (d - r).abs()
print("border err mean", torch.cat([e[..., 0, :].flatten(), e[..., -1, :].flatten(), e[..., :, 0].flatten(), e[..., :, -1].flatten()]).mean().item(),
      "interior", e[..., 1:-1, 1:-1].mean().item(), "ref abs mean", r.abs().mean().item())
for k, p in value.state_dict(keep_vars=True).items():
    if "conv1" in k and "blocks" not in k:
        print(k, ((p.grad - sd[k].grad).norm() / sd[k].grad.norm()).item())
# ratio check: is our dx a scaled version?
print("dot ratio", (d * r).sum().item() / (r * r).sum().item())
# torch's own bf16 regime (autocast: bf16 activations / gradients, fp32 accumulation) against its fp32 result
sd2 = {k: v.cuda().requires_grad_(True) for k, v in vsd.items()}
xa = x.cuda().requires_grad_(True)
with torch.autocast("cuda", dtype=torch.bfloat16):
    oa = nets.value_forward(sd2, xa)
(oa.float().view(-1) * coef).sum().backward()
print("torch autocast-bf16 vs fp32: dx rel", ((xa.grad - r).norm() / r.norm()).item(),
      "conv1.weight", ((sd2["net.conv1.weight"].grad - sd["net.conv1.weight"].grad).norm() / sd["net.conv1.weight"].grad.norm()).item(),
      "blocks.0.conv1.weight", ((sd2["net.blocks.0.conv1.weight"].grad - sd["net.blocks.0.conv1.weight"].grad).norm() / sd["net.blocks.0.conv1.weight"].grad.norm()).item())
print("ours vs torch autocast: dx rel", ((d - xa.grad).norm() / xa.grad.norm()).item())
