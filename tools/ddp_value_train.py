"""Training config on N GPUs (BASELINE.json configs[3]: "DDP all-reduce"): the drop-in value net wrapped in
torch.nn.parallel.DistributedDataParallel, one energy-update step (trainer.py:244-264) per rank on its shard.  Checks that
the NCCL-all-reduced gradients DDP leaves in p.grad equal the average of the ranks' local B200-path gradients, and that all
replicas hold identical weights after the Adam step.  Launch: torchrun --nproc-per-node N tools/ddp_value_train.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.nn.parallel import DistributedDataParallel as DDP  # noqa: E402

from common import VALUE_CFG, load_synth_into  # noqa: E402

from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2  # noqa: E402
from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))


def build():
    v = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    load_synth_into(v, seed=1)
    return v.to(dev)


def d_loss(out, half):
    pos, neg = out[:half], out[half:]
    return pos.mean() - neg.mean() + 0.1 * ((pos**2).mean() + (neg**2).mean())


B = 16  # per rank: 8 "real" + 8 generated
g = torch.Generator().manual_seed(100 + rank)
x = torch.randn(B, 3, 32, 32, generator=g).to(dev)
ddp = DDP(build(), device_ids=[dev.index])
local = build()
opt = torch.optim.Adam(ddp.parameters(), lr=1e-4)
loss = d_loss(ddp(x, 10), B // 2)
loss.backward()
d_loss(local(x, 10), B // 2).backward()
worst = 0.0
for (k, p), q in zip(ddp.module.named_parameters(), local.parameters()):
    avg = q.grad.clone()
    dist.all_reduce(avg)
    avg /= world
    err = ((p.grad - avg).norm() / avg.norm().clamp_min(1e-30)).item()
    worst = max(worst, err)
opt.step()
# replicas stay in sync after the step
chk = torch.stack([p.detach().double().sum() for p in ddp.parameters()]).sum()
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN)
dist.all_reduce(hi, op=dist.ReduceOp.MAX)
with torch.no_grad():
    after = ddp.module(x, 10)
ok = worst < 1e-6 and (hi - lo).abs().item() == 0.0 and torch.isfinite(after).all().item()
flag = torch.tensor([1.0 if ok else 0.0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
# ---- sampler update (trainer.py:348-389) with the DDPM U-Net under DDP: sample_step with grad, value term, clip, Adam
from common import DDPM_CFG  # noqa: E402
from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model  # noqa: E402
from diffusion_by_maxentirl_b200.models.DxMI.var_sampler import VARSampler  # noqa: E402


def build_sampler():
    net = Model(**dict(DDPM_CFG, dropout=0.0))
    smp = VARSampler(net, n_timesteps=10, sample_shape=[3, 32, 32], trainable_beta="fix_last")
    load_synth_into(net)
    return smp.to(dev)


Bs = 4
state = torch.randn(Bs, 3, 32, 32, generator=g).to(dev)
zn = torch.randn(Bs, 3, 32, 32, generator=g).to(dev)
tt = torch.tensor([1, 4, 8, 9], device=dev)
vfix = build()
for p_ in vfix.parameters():
    p_.requires_grad_(False)


def sampler_loss(smp):
    d = smp.sample_step(state, tt, noise=zn)
    nt = (tt < 9).float()
    run = (d["control"] ** 2).flatten(1).mean(1) / (2 * d["sigma"].flatten() ** 2)
    return (vfix(d["sample"], tt + 1).flatten() + (0.1 * run - 0.01 * d["entropy"].flatten()) * nt).mean()


smp_ddp, smp_loc = build_sampler(), build_sampler()
smp_ddp.net = DDP(smp_ddp.net, device_ids=[dev.index])
smp_ddp.train()
smp_loc.train()
opt_s = torch.optim.Adam(smp_ddp.parameters(), lr=1e-4)
ls = sampler_loss(smp_ddp)
ls.backward()
sampler_loss(smp_loc).backward()
worst_s = 0.0
for (k, p), q in zip(smp_ddp.net.module.named_parameters(), smp_loc.net.parameters()):
    avg = q.grad.clone()
    dist.all_reduce(avg)
    avg /= world
    worst_s = max(worst_s, ((p.grad - avg).norm() / avg.norm().clamp_min(1e-30)).item())
torch.nn.utils.clip_grad_norm_(smp_ddp.parameters(), 0.1)
opt_s.step()
chk_s = torch.stack([p.detach().double().sum() for p in smp_ddp.parameters()]).sum()
lo_s, hi_s = chk_s.clone(), chk_s.clone()
dist.all_reduce(lo_s, op=dist.ReduceOp.MIN)
dist.all_reduce(hi_s, op=dist.ReduceOp.MAX)
ok_s = worst_s < 1e-6 and (hi_s - lo_s).abs().item() == 0.0
flag_s = torch.tensor([1.0 if ok_s else 0.0], device=dev)
dist.all_reduce(flag_s, op=dist.ReduceOp.MIN)
flag = torch.minimum(flag, flag_s)
if rank == 0:
    print(f"DDP sampler update (DDPM U-Net) on {world} GPUs: loss {ls.item():.5f}, worst |ddp grad - mean(local grads)| rel {worst_s:.2e}, "
          f"replica checksum spread {(hi_s - lo_s).abs().item():.1e} -> {'OK' if flag_s.item() == 1.0 else 'FAILED'}")
    print(f"DDP value-net step on {world} GPUs: loss {loss.item():.5f}, worst |ddp grad - mean(local grads)| rel {worst:.2e}, "
          f"replica checksum spread {(hi - lo).abs().item():.1e} -> {'OK' if flag.item() == 1.0 else 'FAILED'}")
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
