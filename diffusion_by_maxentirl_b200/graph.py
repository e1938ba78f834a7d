"""CUDA-graph replay of a whole sampler rollout (+ optional energy evaluation).

The rollout is a static launch sequence (about 1.7 k kernels for the CIFAR T=4 configuration), so once shapes are fixed
the whole `sampler.sample(...)` -> `value(x_T)` call can be captured once and replayed: the launches no longer cost CPU
time and the ~1 us inter-kernel gaps shrink.  The capture goes through the public API - the graph contains exactly the
kernels an eager call launches.

The captured launches read the library's packed bf16 operand buffers, which `dxmi_repack` refreshes IN PLACE, and the learned sigmas are
recomputed from the live `log_betas` by captured torch ops - so a graph stays valid across optimizer steps as long as the parameter
storages do not move: `live=True` re-checks the parameters' version counters before every replay (re-packing when an optimizer step
happened) and re-captures if a storage moved (`load_state_dict(assign=True)`, `.to()`, `.half()`).  With `live=False` (default,
inference / evaluation) re-create the GraphedRollout after changing weights."""
import torch


class GraphedRollout:
    def __init__(self, sampler, n_sample, device, value=None, labels=None, warmup=2, packed=None, capture_gather=False, live=False):
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
        self.live, self._warmup = bool(live), warmup
        self._capture()

    def _nets(self):
        nets = [self.sampler.net] + ([self.value] if self.value is not None else [])
        out = []
        for n in nets:
            n = n.module if hasattr(n, "module") else n          # DDP
            n = getattr(n, "net", n)                             # TimeIndependentValue -> its encoder
            n = n.module if hasattr(n, "module") else n
            if hasattr(n, "_ensure_handle"):
                out.append(n)
        return out

    def _capture(self):
        warmup = self._warmup
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
        self._sigs = [n._bound_sig for n in self._nets()]

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
        if self.live:
            # optimizer steps bump the parameters' version counters: refresh the packed operands (in place) before the replay
            for n in self._nets():
                n._ensure_handle(self.device)
            if [n._bound_sig for n in self._nets()] != self._sigs:  # a storage moved: the plans were rebuilt, the graph is stale
                self._capture()
        if noise is None:
            self.noise.normal_()
        else:
            self.noise.copy_(noise, non_blocking=True)
        self.graph.replay()
        return self.out
