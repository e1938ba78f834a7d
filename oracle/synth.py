"""Deterministic synthetic weights and noise, shared by the golden-vector generator and the tests.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Weights are a pure function of (state_dict key, shape, seed), so
the same tensors can be rebuilt anywhere (the GPU box has no /root/reference and a 143 MB state_dict cannot be
committed).  This also removes the reference's zero-initialised tensors (`zero_module`), which would
make a random-init EDM U-Net output exactly 0 and exercise no contraction.
"""
import math
import zlib

import torch


def _gen(key, seed):
    return torch.Generator().manual_seed((zlib.crc32(key.encode()) + 1000003 * seed) % (2**31 - 1))


def synth_tensor(key, shape, seed=0):
    g = _gen(key, seed)
    shape = tuple(shape)
    if key.endswith(".weight") and len(shape) >= 2:
        if key == "label_emb.weight":
            return torch.randn(shape, generator=g) * 0.5
        fan_in = math.prod(shape[1:])
        return torch.randn(shape, generator=g) / math.sqrt(fan_in)
    if key.endswith(".weight"):  # GroupNorm gain
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)  # biases


def synth_state_dict(shapes, seed=0, skip=("log_betas", "std")):
    """shapes: ordered mapping key -> shape.  Returns fp32 tensors for every key not in `skip`."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items() if k not in skip}


def synth_noise(T, B, shape, seed=123):
    """z[0..T]: z[0] is x_0 (one generator, T+1 consecutive draws of [B, C, H, W])."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(B, *shape, generator=g) for _ in range(T + 1)]


def synth_labels(B, num_classes=1000, seed=123):
    g = torch.Generator().manual_seed(seed + 1)
    return torch.randint(0, num_classes, (B,), generator=g)
