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
