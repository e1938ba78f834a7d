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


class PackedRollout:
    """The rank's end-of-rollout payload as ONE contiguous uint8 buffer: [B*C*H*W bytes of u8 samples | B fp32 energies].
    The last transition kernel writes the quantised samples straight into `samples_u8` (sampler.sample(u8_out=...)) and the
    value head writes into `energies` (value(x, t, out=...)), so the post-rollout step (generate_large.py:36-50 + the energy
    statistics) is a single all-gather of this buffer with no quantise / pack kernels in between."""

    def __init__(self, B, sample_shape, device, world=None, with_energy=True):
        self.B, self.shape = int(B), tuple(sample_shape)
        self.n_u8 = self.B * int(torch.tensor(self.shape).prod())
        assert self.n_u8 % 4 == 0
        self.with_energy = with_energy
        self.nbytes = self.n_u8 + (4 * self.B if with_energy else 0)
        self.world = (dist.get_world_size() if dist.is_initialized() else 1) if world is None else world
        self.local = torch.zeros(self.nbytes, dtype=torch.uint8, device=device)
        self.samples_u8 = self.local[: self.n_u8].view(self.B, *self.shape)
        self.energies = self.local[self.n_u8:].view(torch.float32) if with_energy else None
        self.gathered = torch.zeros(self.world, self.nbytes, dtype=torch.uint8, device=device) if self.world > 1 else None

    def all_gather(self, group=None):
        """One collective; returns (samples [world*B, C, H, W] u8, energies [world*B] fp32 or None) in rank order."""
        if self.world == 1:
            return self.samples_u8, self.energies
        dist.all_gather_into_tensor(self.gathered.view(-1), self.local, group=group)
        return self.unpack()

    def unpack(self):
        g = self.gathered
        s = g[:, : self.n_u8].reshape(self.world * self.B, *self.shape)
        e = g[:, self.n_u8:].contiguous().view(torch.float32).reshape(-1) if self.with_energy else None
        return s, e


def gather_packed(samples, energies=None, group=None):
    """Convenience form for un-graphed callers: quantise + pack + ONE all-gather (instead of two)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    u8 = samples if samples.dtype == torch.uint8 else quantize_u8(samples)
    if world == 1:
        return u8, (energies.reshape(-1) if energies is not None else None)
    pk = PackedRollout(u8.shape[0], u8.shape[1:], u8.device, world=world, with_energy=energies is not None)
    pk.samples_u8.copy_(u8)
    if energies is not None:
        pk.energies.copy_(energies.reshape(-1).float())
    return pk.all_gather(group)
