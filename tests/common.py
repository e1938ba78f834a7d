"""Shared builders for the parity tests: the drop-in modules on CUDA with the deterministic synthetic weights of
oracle/synth.py (the same tensors the golden fixtures were generated with)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

DDPM_CFG = dict(resolution=32, in_channels=3, out_ch=3, ch=128, ch_mult=(1, 2, 2, 2), num_res_blocks=2,
                attn_resolutions=[16], dropout=0.1)
VALUE_CFG = dict(in_chan=3, out_chan=1, use_spectral_norm=False, keepdim=False, out_activation="linear",
                 avg_pool_dim=1, learn_out_scale=True, nh=128)


def rel_l2(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def golden(name):
    return np.load(os.path.join(GOLD, name))


def load_synth_into(module, seed=0, skip=("log_betas", "std")):
    from oracle import synth

    sd = module.state_dict()
    new = synth.synth_state_dict({k: tuple(v.shape) for k, v in sd.items()}, seed=seed, skip=skip)
    for k, v in new.items():
        sd[k] = v.to(sd[k].dtype)
    module.load_state_dict(sd)
    return {k: v.detach().cpu().clone() for k, v in module.state_dict().items()}


def build_ddpm(T, device="cuda"):
    from diffusion_by_maxentirl_b200.models.DxMI.unet_small import Model
    from diffusion_by_maxentirl_b200.models.DxMI.var_sampler import VARSampler
    from diffusion_by_maxentirl_b200.models.modules import IGEBMEncoderV2
    from diffusion_by_maxentirl_b200.models.value import TimeIndependentValue

    net = Model(**DDPM_CFG)
    sampler = VARSampler(net, n_timesteps=T, sample_shape=[3, 32, 32], trainable_beta="fix_last")
    value = TimeIndependentValue(IGEBMEncoderV2(**VALUE_CFG))
    sd = load_synth_into(net)
    vsd = load_synth_into(value, seed=1)
    sampler.to(device).eval()
    value.to(device).eval()
    return net, sampler, value, sd, vsd


EDM_SMALL_CFG = dict(sigma_min=0.002, sigma_max=80.0, image_size=32, num_channels=64, num_res_blocks=1, num_heads=4,
                     num_heads_upsample=-1, num_head_channels=64, attention_resolutions="16,8,4", channel_mult="1,2,3,4",
                     dropout=0.0, class_cond=True, use_checkpoint=False, use_scale_shift_norm=True, resblock_updown=True,
                     use_fp16=True, use_new_attention_order=False, learn_sigma=False, weight_schedule="uniform",
                     distillation=False)
# configs/imagenet64/T10.yaml:1-21 and configs/lsun/T4.yaml:1-21
EDM_IN64_CFG = dict(EDM_SMALL_CFG, image_size=64, num_channels=192, num_res_blocks=3, attention_resolutions="32,16,8",
                    channel_mult="")
EDM_LSUN_CFG = dict(EDM_SMALL_CFG, image_size=256, num_channels=256, num_res_blocks=2, attention_resolutions="32,16,8",
                    channel_mult="", class_cond=False, use_scale_shift_norm=False)


def build_edm(cfg, T, device="cuda", fp16=True, **sampler_kw):
    """Drop-in EDM U-Net + OpenAIDiffusion with the synthetic (non-zero) weights of oracle/synth.py.
    Returns (unet, sampler, state_dict on the CPU in fp32)."""
    from diffusion_by_maxentirl_b200.models.cm.script_util import create_model_and_diffusion
    from diffusion_by_maxentirl_b200.models.DxMI.openai_diffusion import OpenAIDiffusion

    unet, diffusion = create_model_and_diffusion(**cfg)
    size = cfg["image_size"]
    sampler = OpenAIDiffusion(unet, diffusion, n_timesteps=T, sample_shape=[3, size, size],
                              class_cond=cfg["class_cond"], num_classes=1000 if cfg["class_cond"] else None,
                              trainable_beta="fix_last", sigma_min=0.002, sigma_max=80.0, **sampler_kw)
    sd32 = load_synth_into(unet, skip=("log_betas",))
    if fp16:
        unet.convert_to_fp16()
    unet.to(device).eval()
    return unet, sampler, sd32


def adm_oracle_kwargs(cfg):
    size = cfg["image_size"]
    mult = {256: (1, 1, 2, 2, 4, 4), 128: (1, 1, 2, 3, 4), 64: (1, 2, 3, 4)}[size] if cfg["channel_mult"] == "" else \
        tuple(int(c) for c in cfg["channel_mult"].split(","))
    return dict(image_size=size, model_channels=cfg["num_channels"], channel_mult=mult,
                num_res_blocks=cfg["num_res_blocks"],
                attention_ds=tuple(size // int(r) for r in cfg["attention_resolutions"].split(",")),
                num_head_channels=cfg["num_head_channels"], use_scale_shift_norm=cfg["use_scale_shift_norm"])
