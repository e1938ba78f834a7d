"""Drop-in for the reference's models/DxMI/openai_diffusion.py::OpenAIDiffusion (:10-127) on the B200 path.

`sample()` is one `dxmi_edm_rollout` call (T x [c_in-scaled U-Net forward + fused EDM ancestral transition]);
`sample_step()` is one U-Net forward plus one `dxmi_edm_step`."""
import ctypes as C

import torch
import torch.nn as nn

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.schedule import EdmSchedule


def _inner(net):
    return net.module if hasattr(net, "module") else net


class OpenAIDiffusion:
    def __init__(self, model, diffusion, n_timesteps, sample_shape, class_cond=False, num_classes=0,
                 trainable_beta=False, sigma_min=0.002, sigma_max=80., stochastic_last=False, rho=7.0):
        self.net = model
        self.diffusion = diffusion
        self.class_cond = class_cond
        self.num_classes = num_classes
        self.sample_shape = sample_shape
        self.n_timesteps = n_timesteps
        self.sigma_max = sigma_max
        s = EdmSchedule(n_timesteps, sigma_min, sigma_max, rho, stochastic_last)
        self.sigmas, self.sigma_down, self.sigma_up = s.sigmas, s.sigma_down, s.sigma_up
        self.trainable_beta = trainable_beta
        if trainable_beta:
            self.net.register_parameter("log_betas", nn.Parameter(torch.log(self.sigma_up.clamp(1e-3))))
        else:
            self.net.register_buffer("log_betas", torch.log(self.sigma_up))

    def get_ancestral_step(self, sigmas):
        sigma_from, sigma_to = sigmas[:-1], sigmas[1:]
        sigma_up = (sigma_to**2 * (sigma_from**2 - sigma_to**2) / sigma_from**2) ** 0.5
        sigma_down = (sigma_to**2 - sigma_up**2) ** 0.5
        return sigma_down, sigma_up

    def train(self):
        self.net.train()

    def eval(self):
        self.net.eval()

    def parameters(self):
        return self.net.parameters()

    # ------------------------------------------------------------------ noise scale actually applied (reference :79-92)
    def _tables(self, device):
        """Device copies of the constant [T] tables (made once per device, outside any CUDA-graph capture)."""
        key = str(device)
        cache = self.__dict__.setdefault("_dev_tables", {})
        if key not in cache:
            T = self.n_timesteps
            steps = torch.arange(T)
            cache[key] = {"sigma_up": self.sigma_up.to(device), "terminal": (steps == T - 1).to(device),
                          "non_terminal3": (steps < T - 3).to(device)}
        return cache[key]

    def _noise_sigma_all(self, device):
        """[T] noise scales of all steps, computed on the device from the live log_betas (capturable, no sync)."""
        tab = self._tables(device)
        sigma_up = tab["sigma_up"]
        if not self.trainable_beta:
            return sigma_up
        sigma = torch.exp(_inner(self.net).log_betas.detach().float().to(device))
        if self.trainable_beta == "fix_last":
            sigma = sigma * ~tab["terminal"] + sigma_up * tab["terminal"]
        elif self.trainable_beta == "fix_last3":
            sigma = sigma * tab["non_terminal3"] + sigma_up * ~tab["non_terminal3"]
        return sigma

    def _noise_sigma(self, indices, device):
        """indices: CPU long tensor [B] -> float tensor [B] ON `device` of the noise scale applied at those steps.
        log_betas is read on the device (no host sync; the call stays CUDA-graph capturable)."""
        sigma_up = self.sigma_up[indices].to(device)
        if not self.trainable_beta:
            return sigma_up
        idx = indices.to(device)
        sigma = torch.exp(_inner(self.net).log_betas.detach().float().to(device)[idx])
        if self.trainable_beta == "fix_last":
            terminal = (indices == self.n_timesteps - 1).to(device)
            sigma = sigma * ~terminal + sigma_up * terminal
        elif self.trainable_beta == "fix_last3":
            non_terminal = (indices < self.n_timesteps - 3).to(device)
            sigma = sigma * non_terminal + sigma_up * (~non_terminal)
        return sigma

    def _sample_step_with_grad(self, x, indices, noise=None, **model_kwargs):
        """The reference's sample_step (openai_diffusion.py:66-99) under autograd, for update_sampler_mixed_precision
        (trainer.py:693-746): F through the B200 U-Net backward plan (models/cm/unet.py `_AdmFunction`), the EDM
        preconditioning / ancestral-step arithmetic and the learned noise scale exp(log_betas) as elementwise torch ops on
        [B, C, H, W] fp32 (they carry the graph to log_betas and to F)."""
        device = x.device
        x = x.detach().float()
        sigma = self.sigmas[indices]
        c_skip, c_out, c_in = [c.float().to(device).reshape(-1, 1, 1, 1) for c in self.diffusion.get_scalings(sigma)]
        rescaled_t = 1000 * 0.25 * torch.log(sigma + 1e-44)
        F = self.net(x, rescaled_t.to(device), x_scale=c_in.reshape(-1), **model_kwargs)
        denoised = c_out * F + c_skip * x
        sig = sigma.float().to(device).reshape(-1, 1, 1, 1)
        d = (x - denoised) / sig
        dt = self.sigma_down[indices].float().to(device).reshape(-1, 1, 1, 1) - sig
        mu = x + d * dt
        sigma_up = self.sigma_up[indices].float().to(device)
        if self.trainable_beta:
            s = torch.exp(_inner(self.net).log_betas[indices.to(device)])
            if self.trainable_beta == "fix_last":
                terminal = (indices == self.n_timesteps - 1).to(device)
                s = s * ~terminal + sigma_up * terminal
            elif self.trainable_beta == "fix_last3":
                non_terminal = (indices < self.n_timesteps - 3).to(device)
                s = s * non_terminal + sigma_up * (~non_terminal)
            sigma_up = s
        z = torch.randn_like(mu) if noise is None else noise.to(device=device, dtype=torch.float32)
        samples = mu + z * sigma_up.reshape(-1, 1, 1, 1)
        return {"sample": samples, "mean": mu, "sigma": sigma_up.clamp(1e-4, None)}

    def sample_step(self, x, indices, noise=None, **model_kwargs):
        device = x.device
        indices = indices.cpu()
        if torch.is_grad_enabled() and _inner(self.net).training:
            return self._sample_step_with_grad(x, indices, noise=noise, **model_kwargs)
        B = x.shape[0]
        x = x.detach().contiguous().float()
        sigma = self.sigmas[indices]
        c_skip, c_out, c_in = self.diffusion.get_scalings(sigma)
        rescaled_t = 1000 * 0.25 * torch.log(sigma + 1e-44)
        F = self.net(x, rescaled_t.to(device), x_scale=c_in.to(device), **model_kwargs)
        s_noise = self._noise_sigma(indices, device).float()
        coef = torch.cat([torch.stack([c_skip, c_out, sigma, self.sigma_down[indices]], dim=1).float().to(device),
                          s_noise[:, None]], dim=1).contiguous()
        z = torch.randn_like(x) if noise is None else noise.to(device=device, dtype=torch.float32).contiguous()
        xn, mu = torch.empty_like(x), torch.empty_like(x)
        L.check(L.lib().dxmi_edm_step(L.ptr(x), L.ptr(F), L.ptr(z), L.ptr(coef), L.ptr(xn), L.ptr(mu), B, x[0].numel(),
                                      L.stream_ptr(x)), "dxmi_edm_step")
        return {"sample": xn, "mean": mu, "sigma": s_noise.clamp(1e-4, None)}

    def sample(self, n_sample, device, i_class=None, enable_grad=False, x0=None, noise=None, u8_out=None):
        """Reference OpenAIDiffusion.sample (:101-127).  `noise` (optional, parity contract): the T per-step z tensors
        ([T, B, C, H, W] or a list); without it, `torch.randn_like` draws are made in the reference's order.
        `u8_out` (optional, not in the reference): uint8 [B, C, H, W] CUDA tensor filled by the last transition kernel with
        `((x + 1) * 127.5).clamp(0, 255)` (generate_large.py:43); also returned as d["sample_u8"]."""
        if enable_grad:
            raise NotImplementedError("enable_grad=True (backward through the rollout) is not built on the B200 path")
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("OpenAIDiffusion.sample runs on CUDA only (no CPU fallback)")
        T, B = self.n_timesteps, int(n_sample)
        shape = tuple(self.sample_shape)
        if self.class_cond:
            if i_class is None:
                i_class = torch.randint(0, self.num_classes, (B,), device=device)
            elif isinstance(i_class, int):
                i_class = torch.tensor([i_class] * B, device=device, dtype=torch.long)
            i_class = i_class.to(device=device, dtype=torch.long).contiguous()
        else:
            i_class = None
        net = _inner(self.net)
        net._check_eval()
        buf = torch.empty(T + 1, B, *shape, device=device)
        buf[0] = (torch.randn(B, *shape, device=device) * self.sigma_max) if x0 is None else x0.to(device)
        if noise is None:
            for i in range(T):
                buf[1 + i] = torch.randn(B, *shape, device=device)
        else:
            nz = (torch.stack(list(noise)) if not torch.is_tensor(noise) else noise).to(device=device, dtype=torch.float32)
            assert nz.shape == (T, B, *shape)
            buf[1:] = nz
        h = net._ensure_handle(device)
        idx = torch.arange(T)
        sigma = self.sigmas[idx]
        c_skip, c_out, c_in = self.diffusion.get_scalings(sigma)
        s_noise = self._noise_sigma_all(device).float().contiguous()
        sched = torch.stack([c_in, 1000 * 0.25 * torch.log(sigma + 1e-44), c_skip, c_out, sigma, self.sigma_down[idx]],
                            dim=1).float().contiguous()
        l_sample = torch.empty(T + 1, B, *shape, device=device)
        mean = torch.empty(T, B, *shape, device=device)
        if u8_out is not None:
            assert u8_out.dtype == torch.uint8 and u8_out.is_contiguous() and u8_out.numel() == B * l_sample[0, 0].numel() and u8_out.is_cuda
        L.check(
            L.lib().dxmi_edm_rollout(h, sched.numpy().ctypes.data_as(C.POINTER(C.c_float)), L.ptr(s_noise), T, L.ptr(buf),
                                     L.ptr(i_class),
                                     L.ptr(l_sample), L.ptr(mean), L.ptr(u8_out), B, L.stream_ptr(device)),
            "dxmi_edm_rollout")
        sig_dev = s_noise.clamp(1e-4, None)
        extra = {"sample_u8": u8_out} if u8_out is not None else {}
        return {**extra, "sample": l_sample[T], "l_sample": [l_sample[i] for i in range(T + 1)], "y": i_class,
                "mean": [mean[i] for i in range(T)], "sigma": [sig_dev[i].repeat(B) for i in range(T)]}
