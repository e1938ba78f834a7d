"""Pins the oracle against the reference itself and writes tests/golden/*.npz.

Run in the build container only (needs /root/reference):  python -m oracle.gen_golden [--edm]
TEST INFRASTRUCTURE - see oracle/__init__.py.

For every config it (1) builds the *reference* modules from /root/reference with the harness shims of
SURVEY.md App. C (numpy-2 float64 fix, tuple ch_mult, host-supplied noise through a patched torch.randn /
randn_like, synthetic non-zero weights), (2) runs the reference's own `sample()` / `forward()`, (3) asserts that the
oracle restatement reproduces every tensor to fp32 round-off, and (4) stores the reference outputs as fixtures.
"""
import argparse
import contextlib
import io
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("DXMI_REFERENCE", "/root/reference")
GOLD = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, ROOT)

from oracle import nets, samplers, synth  # noqa: E402


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def import_reference():
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import models.DxMI.var_sampler as vs  # noqa

    # --- shim 1: numpy>=2 keeps float32 scalars in float32 and the bisection never converges.
    # Emulate numpy<2 value-based promotion: f32 - f32 stays f32, every mixed scalar op after it is float64.
    def _log_cont_noise(t, beta_0, beta_T, T):
        delta_beta = float(np.float32(beta_T) - np.float32(beta_0)) / (T - 1)
        _c = (1.0 - float(beta_0)) / delta_beta
        t_1 = float(t) + 1
        return t_1 * np.log(delta_beta) + vs._log_gamma(_c + 1) - vs._log_gamma(_c - t_1 + 1)

    _orig_bisearch = vs.bisearch

    def bisearch(f, domain, target, eps=1e-8):
        return _orig_bisearch(f, domain, float(target), eps)

    vs._log_cont_noise = _log_cont_noise
    vs.bisearch = bisearch
    return vs


@contextlib.contextmanager
def injected_noise(noise):
    """shim 3: torch.randn / torch.randn_like return the host-supplied tensors in order."""
    it = iter(noise)
    orig_randn, orig_like = torch.randn, torch.randn_like

    def randn(*size, **kw):
        t = next(it)
        return t.clone()

    def randn_like(x, **kw):
        t = next(it)
        assert t.shape == x.shape
        return t.clone()

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like


def load_synth(module, seed=0, skip=("log_betas", "std")):
    sd = module.state_dict()
    new = synth.synth_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed=seed, skip=skip)
    for k, v in new.items():
        sd[k] = v.to(sd[k].dtype)
    module.load_state_dict(sd)
    return {k: v.clone() for k, v in module.state_dict().items()}


DDPM_CFG = dict(resolution=32, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2,
                attn_resolutions=[16], dropout=0.1)  # configs/cifar10/T10.yaml:1-10
VALUE_CFG = dict(in_chan=3, out_chan=1, use_spectral_norm=False, keepdim=False, out_activation="linear",
                 avg_pool_dim=1, learn_out_scale=True, nh=128)  # configs/cifar10/T10.yaml:20-31


def gen_ddpm(T, B):
    vs = import_reference()
    from models.DxMI.unet_small import Model
    from models.modules import IGEBMEncoderV2
    from models.value import TimeIndependentValue

    torch.manual_seed(0)
    net = Model(**DDPM_CFG)
    sampler = vs.VARSampler(net, n_timesteps=T, sample_shape=[3, 32, 32], trainable_beta="fix_last")
    value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    sd = load_synth(net)
    vsd = load_synth(value, seed=1)
    sampler.eval()
    value.eval()

    import json
    with open(os.path.join(GOLD, "ddpm_shapes.json"), "w") as f:  # lets the CPU legs rebuild the synthetic weights
        json.dump({"net": {k: list(v.shape) for k, v in net.state_dict().items() if k not in ("log_betas", "std")},
                   "value": {k: list(v.shape) for k, v in value.state_dict().items()}}, f)

    # ---- schedule: oracle restatement must equal the reference buffers exactly
    sched = samplers.var_schedule(T)
    for name in ("continuous_steps", "Gamma_bar", "x_prev_multiplier", "theta_multiplier", "std"):
        ref = getattr(sampler, name)
        assert torch.equal(ref, sched[name]), (name, ref, sched[name])
    assert np.array_equal(sampler.user_defined_eta, sched["user_defined_eta"])
    assert torch.equal(net.log_betas.detach(), sched["log_betas_init"])

    # ---- rollout on host-supplied noise: reference vs oracle
    noise = synth.synth_noise(T, B, (3, 32, 32))
    t0 = time.time()
    with torch.no_grad(), injected_noise(noise):
        d_ref = sampler.sample(B, device="cpu")
    t_ref = time.time() - t0
    with torch.no_grad():
        energy_ref = value(d_ref["sample"], T)
        fn = lambda x, t: nets.ddpm_unet_forward(sd, x, t)
        d_or = samplers.var_rollout(fn, sched, sd["log_betas"], noise)
        energy_or = nets.value_forward(vsd, d_or["sample"])
    worst = 0.0
    for i in range(T + 1):
        worst = max(worst, rel_l2(d_or["l_sample"][i], d_ref["l_sample"][i]))
    for k in ("mean", "control", "logp"):
        for i in range(T):
            worst = max(worst, rel_l2(d_or[k][i], d_ref[k][i]))
    worst = max(worst, rel_l2(energy_or, energy_ref))
    print(f"[ddpm T={T} B={B}] oracle vs reference worst rel-L2 = {worst:.3e} (reference rollout {t_ref:.2f}s)")
    assert worst < 2e-6, worst

    # ---- sample() == looped sample_step() (reference self-consistency)
    with torch.no_grad():
        x = noise[0].clone()
        for i in range(T):
            with injected_noise([noise[i + 1]]):
                d = sampler.sample_step(x, i)
            o = samplers.var_sample_step(fn, sched, sd["log_betas"], x, torch.full((B,), i, dtype=torch.long), noise[i + 1])
            assert rel_l2(o["sample"], d["sample"]) < 2e-6 and rel_l2(o["logp"], d["logp"]) < 2e-5
            x = d["sample"]
        print(f"   looped sample_step vs sample: {rel_l2(x, d_ref['sample']):.3e}")

    # ---- per-layer activations of one forward for kernel-level unit checks
    with torch.no_grad():
        eps0 = net(noise[0], sched["continuous_steps"][0] * torch.ones(B))
    assert rel_l2(d_or["eps"][0], eps0) < 2e-6
    np.savez_compressed(
        os.path.join(GOLD, f"ddpm_T{T}_B{B}.npz"),
        l_sample=torch.stack(d_ref["l_sample"]).numpy(),
        logp=torch.stack(d_ref["logp"]).numpy(),
        mean_last=d_ref["mean"][-1].numpy(),
        control_first=d_ref["control"][0].numpy(),
        eps_first=eps0.numpy(),
        energy=energy_ref.numpy(),
        continuous_steps=sampler.continuous_steps.numpy(),
        x_prev_multiplier=sampler.x_prev_multiplier.numpy(),
        theta_multiplier=sampler.theta_multiplier.numpy(),
        std=sampler.std.numpy(),
        Gamma_bar=sampler.Gamma_bar.numpy(),
        user_defined_eta=sampler.user_defined_eta,
        log_betas=net.log_betas.detach().numpy(),
        state_dict_keys=np.array(list(net.state_dict().keys())),
        value_state_dict_keys=np.array(list(value.state_dict().keys())),
    )


EDM_CFGS = {
    # configs/imagenet64/T10.yaml
    "in64": dict(diffusion=dict(sigma_min=0.002, sigma_max=80.0, image_size=64, num_channels=192, num_res_blocks=3,
                                num_heads=4, num_heads_upsample=-1, num_head_channels=64,
                                attention_resolutions="32,16,8", channel_mult="", dropout=0.0, class_cond=True,
                                use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=True, use_fp16=True,
                                use_new_attention_order=False, learn_sigma=False, weight_schedule="uniform",
                                distillation=False),
                 sampler=dict(sample_shape=[3, 64, 64], n_timesteps=10, class_cond=True, num_classes=1000,
                              trainable_beta="fix_last", sigma_min=0.002, sigma_max=80.0)),
    # configs/lsun/T4.yaml
    "lsun": dict(diffusion=dict(sigma_min=0.002, sigma_max=80.0, image_size=256, num_channels=256, num_res_blocks=2,
                                num_heads=4, num_heads_upsample=-1, num_head_channels=64,
                                attention_resolutions="32,16,8", channel_mult="", dropout=0.0, class_cond=False,
                                use_checkpoint=False, use_scale_shift_norm=False, resblock_updown=True, use_fp16=True,
                                use_new_attention_order=False, learn_sigma=False, weight_schedule="uniform",
                                distillation=False),
                 sampler=dict(sample_shape=[3, 256, 256], n_timesteps=4, class_cond=False, num_classes=None,
                              trainable_beta="fix_last", sigma_min=0.002, sigma_max=80.0, rho=4.0, stochastic_last=True)),
}


def ddpm_resblock_order(n_levels=4, num_res_blocks=2):
    """ResnetBlock prefixes in execution order (= the order of the reference's nn.Dropout calls in train() mode)."""
    order = [f"down.{l}.block.{b}" for l in range(n_levels) for b in range(num_res_blocks)]
    order += ["mid.block_1", "mid.block_2"]
    order += [f"up.{l}.block.{b}" for l in reversed(range(n_levels)) for b in range(num_res_blocks + 1)]
    return order


def train_dropout_masks(shapes, p, seed):
    """Deterministic scaled keep masks (0 or 1/(1-p)), one per ResnetBlock, NCHW fp32 - shared by the generator and the CPU test."""
    g = torch.Generator().manual_seed(seed)
    return {k: (torch.rand(shp, generator=g) >= p).float() / (1.0 - p) for k, shp in shapes.items()}


def gen_ddpm_train(B=2, p_drop=0.3):
    """Row a9 (training mode): the reference Model in train() mode, its nn.Dropout replaced by host-supplied masks, forward +
    backward of a fixed linear functional of eps; the oracle with the same masks must reproduce eps and every gradient.
    Writes tests/golden/ddpm_train_B2.npz (eps and a few gradient tensors of the reference)."""
    import_reference()
    from models.DxMI.unet_small import Model

    torch.manual_seed(0)
    net = Model(**dict(DDPM_CFG, dropout=p_drop))
    sd = load_synth(net, skip=())
    net.train()
    cm = {"0": 128, "1": 256, "2": 256, "3": 256}
    res = {"0": 32, "1": 16, "2": 8, "3": 4}
    shapes = {}
    for k in ddpm_resblock_order():
        lvl = k.split(".")[1] if not k.startswith("mid") else "3"
        shapes[k] = (B, cm[lvl], res[lvl], res[lvl])
    masks = train_dropout_masks(shapes, p_drop, seed=77)
    g = torch.Generator().manual_seed(78)
    x = torch.randn(B, 3, 32, 32, generator=g)
    t = torch.tensor([394.07648, 66.86534])[:B]
    coef = torch.randn(B, 3, 32, 32, generator=g)
    it = iter([masks[k] for k in ddpm_resblock_order()])
    orig = torch.nn.functional.dropout

    def dropout(inp, p=0.5, training=True, inplace=False):
        assert training and abs(p - p_drop) < 1e-12
        m = next(it)
        assert m.shape == inp.shape, (m.shape, inp.shape)
        return inp * m

    torch.nn.functional.dropout = dropout
    try:
        ref = net(x, t)
    finally:
        torch.nn.functional.dropout = orig
    (ref * coef).sum().backward()
    rsd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out = nets.ddpm_unet_forward(rsd, x, t, dropout_masks=masks)
    (out * coef).sum().backward()
    e = rel_l2(out, ref)
    worst = max((rel_l2(rsd[k].grad, p_.grad), k) for k, p_ in net.named_parameters() if not k.endswith(".k.bias"))
    print(f"[ddpm train B={B} p={p_drop}] oracle vs reference: eps {e:.2e}, worst gradient {worst[1]} {worst[0]:.2e}")
    assert e < 1e-5 and worst[0] < 1e-4
    keep = ["conv_out.weight", "conv_in.weight", "mid.attn_1.q.weight", "down.1.downsample.conv.weight", "up.1.block.0.nin_shortcut.weight",
            "up.2.upsample.conv.weight", "temb.dense.0.weight", "down.0.block.0.temb_proj.weight", "mid.block_1.norm2.weight"]
    grads = dict(net.named_parameters())
    # ---- value net under autograd (trainer.py:244-326): reference vs oracle, asserted here (no fixture: no randomness involved)
    from models.modules import IGEBMEncoderV2
    from models.value import TimeIndependentValue

    value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    vsd = load_synth(value, seed=1)
    xv = torch.randn(4, 3, 32, 32, generator=g).requires_grad_(True)
    cv = torch.randn(4, generator=g)
    (value(xv, 3).flatten() * cv).sum().backward()
    rv = {k: v.clone().requires_grad_(True) for k, v in vsd.items()}
    xo = xv.detach().clone().requires_grad_(True)
    (nets.value_forward(rv, xo).flatten() * cv).sum().backward()
    wv = max(rel_l2(rv[k].grad, p_.grad) for k, p_ in value.named_parameters())
    print(f"[value train] oracle vs reference: worst parameter gradient {wv:.2e}, input gradient {rel_l2(xo.grad, xv.grad):.2e}")
    assert wv < 1e-5 and rel_l2(xo.grad, xv.grad) < 1e-5
    vg = dict(value.named_parameters())
    np.savez_compressed(os.path.join(GOLD, "value_train_B4.npz"), x=xv.detach().numpy(), coef=cv.numpy(), dx=xv.grad.numpy(),
                        **{"grad:" + k: vg[k].grad.numpy()[:8] for k in ("net.conv1.weight", "net.blocks.0.conv2.weight",
                                                                          "net.blocks.2.skip.0.weight", "net.blocks.5.conv1.weight",
                                                                          "net.linear.weight", "net.out_scale.weight")})
    # big tensors: the first 8 output channels only (fixture size)
    np.savez_compressed(os.path.join(GOLD, "ddpm_train_B2.npz"), eps=ref.detach().numpy(), p_drop=np.float64(p_drop),
                        **{"grad:" + k: grads[k].grad.numpy()[:8] for k in keep})


ADM_TRAIN_SMALL = dict(image_size=32, num_channels=64, num_res_blocks=1, channel_mult="1,2,3,4", attention_resolutions="16,8,4")


def gen_adm_train(B=2):
    """Row f4 (EDM training): the reference UNetModel (models/cm/unet.py, reduced width, fp16 torso - the only mode the reference
    supports) in train() mode under autograd: F = net(x, t, y), backward of a fixed linear functional.  The fp16-torso
    oracle with torch autograd must reproduce F and every gradient; the fully-fp32 oracle - the yardstick of the CUDA backward
    tests - must agree within the fp16 noise.
    Writes tests/golden/adm_train_B2.npz (F and a few gradient tensors of the reference)."""
    import_reference()
    from models.cm.script_util import create_model_and_diffusion

    dcfg = dict(EDM_CFGS["in64"]["diffusion"])
    dcfg.update(ADM_TRAIN_SMALL)
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        unet, _ = create_model_and_diffusion(**dcfg)
    sd32 = load_synth(unet, skip=())
    unet.convert_to_fp16()
    unet.train()
    sd16 = {k: v.detach().clone() for k, v in unet.state_dict().items()}
    g = torch.Generator().manual_seed(31)
    x = torch.randn(B, 3, 32, 32, generator=g)
    t = torch.tensor([1.0954, -0.8327])[:B]
    y = torch.tensor([17, 803])[:B]
    coef = torch.randn(B, 3, 32, 32, generator=g)
    ref = unet(x, t, y)
    (ref * coef).sum().backward()
    akw = dict(image_size=32, model_channels=64, channel_mult=(1, 2, 3, 4), num_res_blocks=1, attention_ds=(2, 4, 8),
               num_head_channels=64, use_scale_shift_norm=True)

    def oracle(fp16):
        src = sd16 if fp16 else sd32
        rsd = {k: (v[..., None] if v.dim() == 3 else v).clone().requires_grad_(True) for k, v in src.items()}
        out = nets.adm_unet_forward(rsd, x, t, y, fp16_torso=fp16, **akw)
        (out * coef).sum().backward()
        return out, rsd

    named = dict(unet.named_parameters())
    for fp16 in (True, False):
        out, rsd = oracle(fp16)
        e = rel_l2(out, ref)
        worst = max((rel_l2(rsd[k].grad.reshape(p_.shape).float(), p_.grad.float()), k) for k, p_ in named.items() if p_.grad is not None)
        print(f"[adm train B={B}] oracle (fp16 torso={fp16}) vs reference: F {e:.2e}, worst gradient {worst[1]} {worst[0]:.2e}")
        assert (e < 1e-3 and worst[0] < 2e-2) if fp16 else (e < 1e-2 and worst[0] < 5e-2)
    keep = ["out.2.weight", "input_blocks.0.0.weight", "input_blocks.3.1.qkv.weight", "input_blocks.3.0.in_layers.2.weight",
            "input_blocks.3.0.skip_connection.weight", "input_blocks.2.0.in_layers.2.weight", "output_blocks.1.2.out_layers.3.weight",
            "time_embed.0.weight", "input_blocks.1.0.emb_layers.1.weight", "middle_block.0.out_layers.0.weight",
            "output_blocks.0.0.skip_connection.weight", "middle_block.1.proj_out.weight"]
    np.savez_compressed(os.path.join(GOLD, "adm_train_B2.npz"), F=ref.detach().numpy(), x=x.numpy(), coef=coef.numpy(),
                        t=t.numpy(), y=y.numpy(), label_rows=named["label_emb.weight"].grad[y].float().numpy(),
                        **{"grad:" + k: named[k].grad.float().numpy()[:8] for k in keep})


def gen_edm(name, B, T=None, small=None, seed=123, stride=1, skip_fp32=False):
    import_reference()
    from models.cm.script_util import create_model_and_diffusion
    from models.DxMI.openai_diffusion import OpenAIDiffusion

    cfg = EDM_CFGS[name]
    dcfg = dict(cfg["diffusion"])
    scfg = dict(cfg["sampler"])
    if T is not None:
        scfg["n_timesteps"] = T
    T = scfg["n_timesteps"]
    if small:  # reduced-width variant with the same block structure, for fast CPU/GPU parity tests
        dcfg.update(small)
        scfg["sample_shape"] = [3, dcfg["image_size"], dcfg["image_size"]]
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        unet, diffusion = create_model_and_diffusion(**dcfg)
        sampler = OpenAIDiffusion(unet, diffusion, **scfg)
    sd32 = load_synth(unet, skip=("log_betas",))
    unet.convert_to_fp16()
    unet.eval()
    sd16 = {k: v.clone() for k, v in unet.state_dict().items()}

    sched = samplers.edm_schedule(T, scfg["sigma_min"], scfg["sigma_max"], rho=scfg.get("rho", 7.0),
                                  stochastic_last=scfg.get("stochastic_last", False))
    assert torch.equal(sched["sigmas"], sampler.sigmas)
    assert torch.equal(sched["sigma_up"], sampler.sigma_up) and torch.equal(sched["sigma_down"], sampler.sigma_down)
    assert torch.equal(sched["log_betas_init"], unet.log_betas.detach())

    shape = tuple(scfg["sample_shape"])
    noise = synth.synth_noise(T, B, shape, seed=seed)
    noise[0] = noise[0] * scfg["sigma_max"]
    y = synth.synth_labels(B, seed=seed) if scfg.get("class_cond") else None
    t0 = time.time()
    with torch.no_grad(), injected_noise(noise[1:]):
        d_ref = sampler.sample(B, device="cpu", i_class=y, x0=noise[0])
    t_ref = time.time() - t0

    size = dcfg["image_size"]
    from models.cm.script_util import create_model  # noqa: F401  (channel_mult resolution mirrors create_model)
    mult = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[size] if dcfg["channel_mult"] == "" else \
        tuple(int(c) for c in dcfg["channel_mult"].split(","))
    akw = dict(image_size=size, model_channels=dcfg["num_channels"], channel_mult=mult,
               num_res_blocks=dcfg["num_res_blocks"],
               attention_ds=tuple(size // int(r) for r in dcfg["attention_resolutions"].split(",")),
               num_head_channels=dcfg["num_head_channels"], use_scale_shift_norm=dcfg["use_scale_shift_norm"])
    with torch.no_grad():
        fn16 = lambda x, t, yy: nets.adm_unet_forward(sd16, x, t, yy, fp16_torso=True, **akw)
        d_or = samplers.edm_rollout(fn16, sched, sd16["log_betas"], noise, y)
    worst = max(rel_l2(d_or["l_sample"][i], d_ref["l_sample"][i]) for i in range(T + 1))
    print(f"[edm {name} T={T} B={B} small={bool(small)}] oracle(fp16 torso) vs reference worst rel-L2 = {worst:.3e} "
          f"(reference rollout {t_ref:.1f}s)")
    assert worst < 1e-5, worst
    if skip_fp32:
        d32 = d_or
    else:
        with torch.no_grad():
            fn32 = lambda x, t, yy: nets.adm_unet_forward(sd32, x, t, yy, fp16_torso=False, **akw)
            d32 = samplers.edm_rollout(fn32, sched, sd32["log_betas"], noise, y)
        drift = [rel_l2(d_ref["l_sample"][i], d32["l_sample"][i]) for i in range(T + 1)]
        print("   reference fp16-torso vs oracle fp32, per state:", " ".join(f"{v:.1e}" for v in drift))
    tag = f"edm_{name}{'_small' if small else ''}_T{T}_B{B}"
    # stride > 1: a strided pixel slice of every state (rel-L2 over the slice is the test statistic; keeps fixtures small)
    sl = (slice(None), slice(None), slice(None), slice(None, None, stride), slice(None, None, stride))
    np.savez_compressed(
        os.path.join(GOLD, tag + ".npz"),
        l_sample_ref_fp16=torch.stack(d_ref["l_sample"]).numpy()[sl],
        l_sample_fp32=torch.stack(d32["l_sample"]).numpy()[sl],
        mean_ref_fp16=torch.stack(d_ref["mean"]).numpy()[sl],
        stride=np.int64(stride), seed=np.int64(seed),
        F_first_fp32=d32["F"][0].numpy()[sl[1:]],
        sigmas=sampler.sigmas.numpy(), sigma_up=sampler.sigma_up.numpy(), sigma_down=sampler.sigma_down.numpy(),
        log_betas=unet.log_betas.detach().numpy(),
        y=(y.numpy() if y is not None else np.zeros(0, dtype=np.int64)),
        state_dict_keys=np.array(list(unet.state_dict().keys())),
        state_dict_dtypes=np.array([str(v.dtype) for v in unet.state_dict().values()]),
    )


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--edm", action="store_true")
    ap.add_argument("--edm-full", action="store_true")
    ap.add_argument("--edm-full-t10", action="store_true")
    ap.add_argument("--lsun", action="store_true")
    ap.add_argument("--skip-ddpm", action="store_true")
    ap.add_argument("--train", action="store_true", help="row a9: pin the training-mode (dropout + autograd) oracle")
    ap.add_argument("--adm-train", action="store_true", help="row f4: pin the ADM U-Net under autograd")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    if not args.skip_ddpm:
        gen_ddpm(T=10, B=2)
        gen_ddpm(T=4, B=2)
    if args.train:
        gen_ddpm_train()
    if args.adm_train:
        gen_adm_train()
    if args.edm:
        gen_edm("in64", B=2, T=4, small=dict(image_size=32, num_channels=64, num_res_blocks=1,
                                             channel_mult="1,2,3,4", attention_resolutions="16,8,4"))
    if args.edm_full:
        gen_edm("in64", B=1, T=2)
    if args.edm_full_t10:
        # full-width ImageNet-64, the north-star T=10 rollout, B=2, from the reference itself (fp16 torso on the CPU)
        gen_edm("in64", B=2, T=10, seed=21, stride=2)
    if args.lsun:
        # full LSUN-256 T=4 (rho=4, stochastic last step) rollout at B=1 from the reference itself; every 4th pixel kept
        gen_edm("lsun", B=1, T=4, seed=23, stride=4, skip_fp32=True)
