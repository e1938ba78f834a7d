"""CUDA-graph replay of a whole sampler rollout (+ optional energy evaluation).

The rollout is a static launch sequence (about 1.7 k kernels for the CIFAR T=4 configuration), so once shapes are fixed
the whole `sampler.sample(...)` -> `value(x_T)` call can be captured once and replayed: the launches no longer cost CPU
time and the ~1 us inter-kernel gaps shrink.  The capture goes through the public API - the graph contains exactly the
kernels an eager call launches.

The packed bf16 weights are baked into the captured launches: re-create the GraphedRollout after an optimizer step or a
`load_state_dict` (inference / evaluation use)."""
import torch


class GraphedRollout:
    def __init__(self, sampler, n_sample, device, value=None, labels=None, warmup=2, packed=None, capture_gather=False):
        """packed: a dist.PackedRollout - the last transition kernel then writes the u8 samples and the value head the energies
        straight into its buffer; capture_gather: also capture the single NCCL all-gather of that buffer inside the graph."""
        self.sampler, self.value = sampler, value
        self.packed, self.capture_gather = packed, capture_gather and packed is not None and packed.world > 1
        self.B, self.T = int(n_sample), sampler.n_timesteps
        self.device = torch.device(device)
        shape = tuple(sampler.sample_shape)
        self.edm = hasattr(sampler, "sigma_max")
        self.noise = torch.zeros(self.T + 1, self.B, *shape, device=self.device)  # noise[0] = x_0 (unit variance)
        self.labels = labels.to(self.device).clone() if labels is not None else None
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                self._run()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = self._run()

    @torch.no_grad()
    def _run(self):
        u8 = self.packed.samples_u8 if self.packed is not None else None
        if self.edm:
            d = self.sampler.sample(self.B, self.device, i_class=self.labels, x0=self.noise[0] * self.sampler.sigma_max,
                                    noise=self.noise[1:], u8_out=u8)
        else:
            d = self.sampler.sample(self.B, device=self.device, noise=self.noise, u8_out=u8)
        e = None
        if self.value is not None:
            if self.packed is not None and self.packed.with_energy:
                e = self.value(d["sample"], self.T, out=self.packed.energies)
            else:
                e = self.value(d["sample"], self.T)
        if self.capture_gather:
            self.packed.all_gather()
        return d, e

    def __call__(self, noise=None):
        """noise: [T+1, B, C, H, W] (device or pinned host; noise[0] is x_0 before the sigma_max scaling for EDM) or None to
        draw fresh Gaussian noise.  Returns (d_sample dict, energies) - views of static buffers, overwritten by the next call."""
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise, non_blocking=True)
        self.graph.replay()
        return self.out
