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
