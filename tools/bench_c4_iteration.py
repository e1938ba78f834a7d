"""BASELINE.json configs[3]: one DxMI training iteration of the CIFAR-10 DDPM T=10 configuration at batch 128 per GPU on the
B200 path, following the schedule of SURVEY 3.3 (train_cifar10.py:141-205, trainer.py:230-408):
  rollout (sampler.eval(), no grad, T U-Net forwards)  ->  energy update of the value net on cat(real, x_T)  ->  T TD updates of the
  value net (target v(next, t+1) without grad, mse on v(state, t), clip 0.1, Adam)  ->  sampler update (sampler.train(), dropout
  0.1, sample_step with grad on B random buffer rows, value term + running cost - entropy, clip 0.1, Adam).
This is a timing harness over the public drop-in modules (synthetic "real" images, random-init weights); the losses restate the
trainer's structure, not its exact hyper-parameters.  Usage: python tools/bench_c4_iteration.py [B] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from common import DDPM_CFG, VALUE_CFG, load_synth_into  # noqa: E402

from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model  # noqa: E402
from diffusion_by_maxentirl_b200.models.DxMI.var_sampler import VARSampler  # noqa: E402
from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2  # noqa: E402
from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = 10
dev = "cuda"
net = Model(**DDPM_CFG)  # dropout 0.1 as in configs/cifar10/T10.yaml
sampler = VARSampler(net, n_timesteps=T, sample_shape=[3, 32, 32], trainable_beta="fix_last")
load_synth_into(net)
sampler.to(dev)
v = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
load_synth_into(v, seed=1)
v.to(dev)
opt_v = torch.optim.Adam(v.parameters(), lr=1e-5)
opt_s = torch.optim.Adam(sampler.parameters(), lr=1e-6)
images = torch.rand(B, 3, 32, 32, device=dev) * 2 - 1
ev = {k: torch.cuda.Event(enable_timing=True) for k in ("a", "b", "c", "d", "e")}


def iteration(timed):
    if timed:
        ev["a"].record()
    sampler.eval()
    with torch.no_grad():
        d = sampler.sample(B, device=dev)
    if timed:
        ev["b"].record()
    xs = torch.stack(d["l_sample"])  # [T+1, B, ...]
    # ---- energy update (trainer.py:244-264)
    out = v(torch.cat([images, xs[-1]]), T)
    pos, neg = out[:B], out[B:]
    d_loss = pos.mean() - neg.mean() + 0.05 * ((pos**2).mean() + (neg**2).mean())
    opt_v.zero_grad(set_to_none=True)
    d_loss.backward()
    opt_v.step()
    if timed:
        ev["c"].record()
    # ---- T TD updates (trainer.py:276-326)
    for i in range(T):
        tt = T - 1 - i
        state, nxt = xs[tt], xs[tt + 1]
        with torch.no_grad():
            ctrl = d["control"][tt]
            running = (ctrl**2).flatten(1).mean(1) / (2 * d["sigma"][tt].flatten() ** 2)
            target = v(nxt, tt + 1).flatten() + 0.1 * running
        v_loss = F.mse_loss(v(state, tt).flatten(), target)
        opt_v.zero_grad(set_to_none=True)
        v_loss.backward()
        torch.nn.utils.clip_grad_norm_(v.parameters(), 0.1)
        opt_v.step()
    if timed:
        ev["d"].record()
    # ---- sampler update (trainer.py:348-389)
    sampler.train()
    idx_t = torch.randint(0, T, (B,), device=dev)
    state = xs[idx_t, torch.arange(B, device=dev)]
    ds = sampler.sample_step(state, idx_t)
    nt = (idx_t < T - 1).float()
    running = (ds["control"] ** 2).flatten(1).mean(1) / (2 * ds["sigma"].flatten() ** 2)
    for p in v.parameters():
        p.requires_grad_(False)
    s_loss = (v(ds["sample"], idx_t + 1).flatten() + (0.1 * running - 0.01 * ds["entropy"].flatten()) * nt).mean()
    for p in v.parameters():
        p.requires_grad_(True)
    opt_s.zero_grad(set_to_none=True)
    s_loss.backward()
    torch.nn.utils.clip_grad_norm_(sampler.parameters(), 0.1)
    opt_s.step()
    if timed:
        ev["e"].record()
    return d_loss.item(), v_loss.item(), s_loss.item()


iteration(False)
iteration(False)
torch.cuda.synchronize()
acc = [0.0] * 4
for _ in range(iters):
    losses = iteration(True)
    torch.cuda.synchronize()
    for j, (x, y) in enumerate((("a", "b"), ("b", "c"), ("c", "d"), ("d", "e"))):
        acc[j] += ev[x].elapsed_time(ev[y]) / iters
tot = sum(acc)
print(f"C4 iteration B={B} T={T}: {tot:.1f} ms = rollout {acc[0]:.1f} + energy update {acc[1]:.1f} + {T} TD updates {acc[2]:.1f} + sampler update {acc[3]:.1f}")
print(f"  -> {B / tot * 1e3:.0f} generated images / s / GPU through the whole training iteration; ~30.8 TFLOP/iteration at B=128 "
      f"(SURVEY 3.3) -> {30.8 * B / 128 / tot * 1e3:.0f} TFLOP/s; last losses d={losses[0]:.4f} v={losses[1]:.4f} s={losses[2]:.4f}")
