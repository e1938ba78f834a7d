"""the EDM sampler update differentiates F = net(c_in x, c_noise, y) through the ADM U-Net
(trainer.py:693-746, models/cm/unet.py:761-790 under autograd) and steps through MixedPrecisionTrainer
(models/cm/fp16_util.py:161-248).  Every parameter gradient of the B200 backward plan (engine_train_adm.cu) against fp32 CPU
autograd over the oracle; bf16 tolerance 3e-2 on every tensor (the DDPM plan measures 2.3e-2 at the same bar)."""
import statistics

import pytest
import torch

from common import EDM_SMALL_CFG, adm_oracle_kwargs, build_edm, rel_l2

pytestmark = pytest.mark.gpu

# width 128 / 256 at 32x32 / 16x16, attention at both (sequence 1024: per-head materialised backward; 256)
CFG_A = dict(EDM_SMALL_CFG, image_size=32, num_channels=128, channel_mult="1,2", attention_resolutions="32,16", num_res_blocks=1)
# width 192 / 384 (not multiples of 128: overlapping weight-gradient slices) at 16x16 / 8x8, attention at 256 and 64 tokens
CFG_B = dict(EDM_SMALL_CFG, image_size=16, num_channels=192, channel_mult="1,2", attention_resolutions="16,8", num_res_blocks=2)


def _grads_vs_oracle(cfg, B, seed):
    from oracle import nets

    unet, sampler, sd = build_edm(cfg, T=4, fp16=False)
    unet.train()
    size = cfg["image_size"]
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, 3, size, size, generator=g)
    t = torch.randn(B, generator=g) * 1.2 - 0.4   # c_noise = 0.25 ln sigma
    y = torch.randint(0, 1000, (B,), generator=g)
    coef = torch.randn(B, 3, size, size, generator=g)
    # qkv / proj_out are Conv1d in the reference ([3C, C, 1]); the functional oracle applies them as 1x1 conv2d
    rsd = {k: (v[..., None] if v.dim() == 3 else v).clone().requires_grad_(True) for k, v in sd.items() if k != "log_betas"}
    ref = nets.adm_unet_forward(rsd, x, t, y, fp16_torso=False, **adm_oracle_kwargs(cfg))
    (ref * coef).sum().backward()
    out = unet(x.cuda(), t.cuda(), y.cuda())
    assert out.requires_grad
    assert rel_l2(out, ref) < 2e-2, rel_l2(out, ref)
    (out * coef.cuda()).sum().backward()
    torch.cuda.synchronize()
    errs = {}
    for k, p in unet.named_parameters():
        if k == "log_betas":
            continue
        assert p.grad is not None, k
        r = rsd[k].grad.reshape(p.shape)
        if k == "label_emb.weight":
            # only the rows of the batch's labels are touched
            rows = torch.unique(y)
            assert torch.count_nonzero(p.grad.cpu()).item() <= rows.numel() * p.shape[1]
            errs[k] = rel_l2(p.grad.cpu()[rows], r[rows])
            continue
        if k.endswith(".qkv.bias"):
            # softmax is invariant to a constant added to every key's score: the k third of this gradient is exactly zero in
            # exact arithmetic - compare the q and v thirds
            C = p.shape[0] // 3
            errs[k] = max(rel_l2(p.grad[:C], r[:C]), rel_l2(p.grad[2 * C:], r[2 * C:]))
            assert p.grad[C:2 * C].abs().max().item() < 2e-2 * max(r[:C].abs().max().item(), 1e-6), k
            continue
        errs[k] = rel_l2(p.grad, r)
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:10]
    print("worst ADM gradient errors:", [(k, "%.2e" % e) for k, e in worst])
    print("median %.2e over %d tensors" % (statistics.median(errs.values()), len(errs)))
    return errs


def test_adm_backward_matches_autograd_128():
    errs = _grads_vs_oracle(CFG_A, B=2, seed=5)
    for k, e in errs.items():
        assert e < 3e-2, (k, e)


def test_adm_backward_matches_autograd_192():
    errs = _grads_vs_oracle(CFG_B, B=3, seed=6)
    for k, e in errs.items():
        assert e < 3e-2, (k, e)


def test_mixed_precision_trainer_step():
    """convert_to_fp16 + MixedPrecisionTrainer(use_fp16=True, special_key='log_betas') as train_image_large.py:155-163 builds it:
    loss-scaled backward through the B200 plan, unscale, optimizer step on the fp32 masters, copy back into the fp16 model
    parameters (which re-packs the tensor-core operands: the next forward must change), lg_loss_scale growth; an overflowing
    step is skipped and lowers the exponent."""
    from diffusion_by_maxentirl_b200.models.cm.fp16_util import MixedPrecisionTrainer

    unet, sampler, sd = build_edm(CFG_A, T=4, fp16=True)
    unet.train()
    mp = MixedPrecisionTrainer(model=unet, use_fp16=True, initial_lg_loss_scale=8.0, special_key="log_betas")
    assert len(mp.master_params) == 3 and mp.master_params[0].numel() == unet.log_betas.numel()
    assert all(m.dtype == torch.float32 for m in mp.master_params)
    opt = torch.optim.RAdam([{"params": mp.master_params[1:], "lr": 1e-3}, {"params": mp.master_params[0:1], "lr": 1e-2}])
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 32, 32, generator=g).cuda()
    t = torch.tensor([0.3, -1.1]).cuda()
    y = torch.tensor([7, 901]).cuda()
    with torch.no_grad():
        unet.eval()
        before = unet(x, t, y)
        unet.train()
    mp.zero_grad()
    out = unet(x, t, y)
    loss = (out ** 2).mean() + unet.log_betas.sum() * 0.0
    mp.backward(loss)
    w = unet._param("input_blocks.1.0.in_layers.2.weight")
    assert w.dtype == torch.float16 and w.grad is not None and w.grad.dtype == torch.float16
    w0 = w.detach().clone()
    assert mp.optimize(opt) is True
    assert abs(mp.lg_loss_scale - 8.001) < 1e-9
    assert not torch.equal(w0, w.detach())
    with torch.no_grad():
        unet.eval()
        after = unet(x, t, y)
        unet.train()
    assert 0 < rel_l2(after, before) < 0.5
    # overflow: a huge loss scale makes the fp16 gradients inf -> the step is skipped, the exponent drops by one
    mp.lg_loss_scale = 60.0
    mp.zero_grad()
    out = unet(x, t, y)
    mp.backward((out ** 2).mean() * 1e6)
    w1 = w.detach().clone()
    assert mp.optimize(opt) is False
    assert mp.lg_loss_scale == 59.0
    assert torch.equal(w1, w.detach())


def test_edm_sampler_update_like_the_trainer():
    """trainer.py:693-746 (update_sampler_mixed_precision): d = sampler.sample_step(state, t, y=y) with grad; loss = mean(v(next,
    t+1) + (tau2 running_cost - tau1 log sigma) non_terminal); mp_trainer.backward / optimize.  Loss, the log_betas gradient
    and a torso convolution gradient against the same step in fp32 autograd over the oracle on the CPU."""
    from test_train_gpu import build_value
    from diffusion_by_maxentirl_b200.models.cm.fp16_util import MixedPrecisionTrainer
    from oracle import nets, samplers

    cfg = CFG_A
    T = 4
    unet, sampler, sd = build_edm(cfg, T=T, fp16=True)
    value, vsd = build_value()
    for p in value.parameters():
        p.requires_grad_(False)
    sampler.train()
    mp = MixedPrecisionTrainer(model=unet, use_fp16=True, initial_lg_loss_scale=0.0, special_key="log_betas")  # (synthetic weights at sigma = 80 give O(100) gradients: larger scales overflow fp16, which test_mixed_precision_trainer_step covers)
    opt = torch.optim.RAdam([{"params": mp.master_params[1:], "lr": 1e-8}, {"params": mp.master_params[0:1], "lr": 1e-6}])
    B = 4
    g = torch.Generator().manual_seed(21)
    sched = samplers.edm_schedule(T)
    t = torch.tensor([0, 1, 2, 3])
    state = torch.randn(B, 3, 32, 32, generator=g) * sched["sigmas"][t].float()[:, None, None, None]
    z = torch.randn(B, 3, 32, 32, generator=g)
    y = torch.tensor([3, 500, 77, 999])
    tau1, tau2, skip_tau = 0.1, 0.01, 1
    betas_q = torch.exp(unet.log_betas.detach().cpu()) ** 2  # use_sampler_beta: betas_for_q from the sampler's sigmas

    def loss_fn(d, st, v_next, tt):
        beta_next = betas_q[(T - tt.cpu() - 1)].to(st.device)
        running = ((d["sample"] - st) ** 2).flatten(1).mean(1) / (2 * beta_next)
        non_terminal = (tt < T - skip_tau).float().to(st.device)
        return (v_next.flatten() + (running * tau2 - torch.log(d["sigma"].flatten()) * tau1) * non_terminal).mean()

    mp.zero_grad()
    d = sampler.sample_step(state.cuda(), t.cuda(), noise=z.cuda(), y=y.cuda())
    assert d["sample"].requires_grad
    loss = loss_fn(d, state.cuda(), value(d["sample"], t.cuda() + 1), t)
    mp.backward(loss)
    scale = 2 ** mp.lg_loss_scale
    lb_grad = unet.log_betas.grad.detach().float().cpu() / scale
    wk = "input_blocks.1.0.in_layers.2.weight"
    w_grad = unet._param(wk).grad.detach().float().cpu() / scale
    assert mp.optimize(opt) is True
    # ---- oracle: the same step, fp32 autograd on the CPU
    rsd = {k: (v[..., None] if v.dim() == 3 else v).clone().requires_grad_(True) for k, v in sd.items() if k != "log_betas"}
    lb = sched["log_betas_init"].clone().requires_grad_(True)
    sigma = sched["sigmas"][t].float()
    c_skip, c_out, c_in = [c.float()[:, None, None, None] for c in samplers.edm_scalings(sigma)]
    F = nets.adm_unet_forward(rsd, c_in * state, 1000 * 0.25 * torch.log(sigma + 1e-44), y, fp16_torso=False, **adm_oracle_kwargs(cfg))
    den = c_out * F + c_skip * state
    sig = sigma[:, None, None, None]
    mu = state + (state - den) / sig * (sched["sigma_down"][t].float()[:, None, None, None] - sig)
    s_up = torch.exp(lb[t])
    terminal = t == T - 1
    s_up = s_up * ~terminal + sched["sigma_up"][t].float() * terminal
    xn = mu + z * s_up[:, None, None, None]
    rvsd = {k: v.clone() for k, v in vsd.items()}
    rloss = loss_fn({"sample": xn, "sigma": s_up.clamp(1e-4, None)}, state, nets.value_forward(rvsd, xn), t)
    rloss.backward()
    print(f"EDM sampler loss {loss.item():.6f} vs oracle {rloss.item():.6f}")
    assert abs(loss.item() - rloss.item()) < 2e-2 * max(1.0, abs(rloss.item()))
    e_lb = rel_l2(lb_grad[:-1], lb.grad[:-1])
    e_w = rel_l2(w_grad, rsd[wk].grad)
    print(f"log_betas grad rel-L2 {e_lb:.2e}; {wk} grad rel-L2 {e_w:.2e}")
    assert e_lb < 5e-2 and e_w < 5e-2
