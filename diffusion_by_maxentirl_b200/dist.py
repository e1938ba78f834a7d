"""Multi-GPU plumbing of the rollout path: one process per GPU (torchrun), batch sharding, and the path's only exchange -
the end-of-rollout all-gather of uint8 samples (generate_large.py:36-50) and fp32 energies.

The path shards by batch with no data-path collective (every image's trajectory is independent given its noise, SURVEY
8e); NCCL is only used for this gather (and by PyTorch DDP for gradients in the training configs)."""
import torch
import torch.distributed as dist

from . import _lib as L


def shard_bounds(n_total, rank, world):
    """[lo, hi) of the contiguous slice of a global batch owned by `rank` (first n_total % world ranks get one extra)."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_noise(noise, rank, world):
    """Per-rank slice [T+1, B_local, C, H, W] of host-supplied noise [T+1, B_global, C, H, W] (parity contract: the
    union of all ranks' samples equals the single-GPU run on the same noise)."""
    lo, hi = shard_bounds(noise.shape[1], rank, world)
    return noise[:, lo:hi].contiguous()


def quantize_u8(samples):
    """fp32 samples in [-1, 1] -> uint8 via ((x + 1) * 127.5).clamp(0, 255) (generate_large.py:43) on the GPU."""
    if samples.device.type != "cuda":
        raise RuntimeError("quantize_u8 runs on CUDA only (no CPU fallback)")
    x = samples.detach().contiguous().float()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    L.check(L.lib().dxmi_quantize_u8(L.ptr(x), L.ptr(out), x.numel(), L.stream_ptr(x)), "dxmi_quantize_u8")
    return out


def gather_rollout(samples_u8, energies=None, group=None):
    """All-gather of equally sized per-rank uint8 samples [B, C, H, W] (+ fp32 energies [B] / [B, 1]) in rank order.
    Returns (samples [world*B, C, H, W], energies [world*B] or None).  Works on any backend (NCCL on GPUs; gloo in the
    CPU tests)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return samples_u8, (energies.reshape(-1) if energies is not None else None)
    s = samples_u8.contiguous()
    out_s = torch.empty((world * s.shape[0],) + tuple(s.shape[1:]), dtype=s.dtype, device=s.device)
    dist.all_gather_into_tensor(out_s, s, group=group)
    out_e = None
    if energies is not None:
        e = energies.reshape(-1).contiguous().float()
        out_e = torch.empty(world * e.shape[0], dtype=torch.float32, device=e.device)
        dist.all_gather_into_tensor(out_e, e, group=group)
    return out_s, out_e
