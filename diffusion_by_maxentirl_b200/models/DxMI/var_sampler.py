"""Drop-in for the reference's models/DxMI/var_sampler.py::VARSampler (:300-444) on the B200 path.

`sample()` is one `dxmi_var_rollout` call (T x [U-Net forward + fused transition kernel]); `sample_step()` is one
U-Net forward plus one `dxmi_var_step`.  Schedule tables come from `diffusion_by_maxentirl_b200.schedule`.
"""
import ctypes as C

import torch
import torch.nn as nn

from diffusion_by_maxentirl_b200 import _lib as L
from diffusion_by_maxentirl_b200.models.modules import process_single_t
from diffusion_by_maxentirl_b200.schedule import VarSchedule


def _inner(net):
    return net.module if hasattr(net, "module") else net


class VARSampler(nn.Module):
    def __init__(self, net, n_timesteps, sample_shape, trainable_beta=True, adhoc_scale1=1.0, adhoc_scale2=1.0):
        super().__init__()
        assert trainable_beta in {True, False, "fix_last"}
        self.net = net
        self.n_timesteps = n_timesteps
        self.sample_shape = sample_shape
        self.adhoc_scale1 = adhoc_scale1
        self.adhoc_scale2 = adhoc_scale2
        self.trainable_beta = trainable_beta
        self.kappa = 1.0
        s = VarSchedule(n_timesteps, self.kappa)
        self.user_defined_eta = s.user_defined_eta
        self.register_buffer("continuous_steps", s.continuous_steps)
        self.register_buffer("Gamma_bar", s.Gamma_bar)
        if trainable_beta:
            _inner(self.net).log_betas = nn.Parameter(torch.log(s.std * adhoc_scale2))  # "log sigma", reference :354-355
        self.register_buffer("x_prev_multiplier", s.x_prev_multiplier)
        self.register_buffer("theta_multiplier", s.theta_multiplier)
        self.register_buffer("std", s.std.clone())
        self.register_buffer("diffusion_steps_list", s.diffusion_steps_list)
        if trainable_beta == "fix_last":
            _inner(self.net).register_buffer("std", s.std.clone())
        # host copy of the constant schedule rows {tau, a, c * adhoc_scale1}: the rollout call needs no device read
        self._sched_host = torch.stack([s.continuous_steps.float(), s.x_prev_multiplier,
                                        s.theta_multiplier * adhoc_scale1], dim=1).contiguous()

    # ------------------------------------------------------------------ per-step noise scale (reference :268-283)
    def _sigmas(self):
        net = _inner(self.net)
        if self.trainable_beta == "fix_last":
            return torch.exp(torch.cat([net.log_betas[:-1], net.std[-1].log().unsqueeze(0)])).detach().float()
        if self.trainable_beta:
            return torch.exp(net.log_betas).detach().float()
        return self.std.float()

    def _forward_net(self, x, t):
        # go through self.net (possibly DDP-wrapped) so hooks / wrappers behave as with the reference
        return self.net(x, t)

    # ------------------------------------------------------------------ rollout
    def sample(self, n_sample, device="cpu", enable_grad=False, noise=None, u8_out=None):
        """Reference VARSampler.sample (:411-428).  `noise` (optional, parity contract): [T+1, B, C, H, W] or a list of
        T+1 tensors; noise[0] is x_0.  Without it, T+1 `torch.randn` draws are made in the reference's order.
        `u8_out` (optional, not in the reference): a uint8 [B, C, H, W] CUDA tensor that the last transition kernel fills with
        the quantised samples (the callers' `((x + 1) * 127.5).clamp(0, 255).to(uint8)`, generate_cifar10.py:205-209);
        also returned as d["sample_u8"]."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("VARSampler.sample runs on CUDA only (no CPU fallback)")
        if enable_grad:
            return self._sample_with_grad(int(n_sample), device, noise)
        T, B = self.n_timesteps, int(n_sample)
        shape = tuple(self.sample_shape)
        net = _inner(self.net)
        net._check_eval()
        if noise is None:
            noise_t = torch.empty(T + 1, B, *shape, device=device)
            for i in range(T + 1):
                noise_t[i] = torch.randn(B, *shape, device=device)
        else:
            noise_t = (torch.stack(list(noise)) if not torch.is_tensor(noise) else noise).to(device=device, dtype=torch.float32).contiguous()
            assert noise_t.shape == (T + 1, B, *shape)
        h = net._ensure_handle(device)
        sig_dev = self._sigmas().to(device).contiguous()  # device tensor: no host sync (and CUDA-graph capturable)
        sched = self._sched_host
        l_sample = torch.empty(T + 1, B, *shape, device=device)
        mean = torch.empty(T, B, *shape, device=device)
        control = torch.empty(T, B, *shape, device=device)
        logp = torch.empty(T, B, device=device)
        if u8_out is not None:
            assert u8_out.dtype == torch.uint8 and u8_out.is_contiguous() and u8_out.numel() == B * l_sample[0, 0].numel() and u8_out.is_cuda
        L.check(
            L.lib().dxmi_var_rollout(h, sched.numpy().ctypes.data_as(C.POINTER(C.c_float)), L.ptr(sig_dev), T, L.ptr(noise_t),
                                     L.ptr(l_sample), L.ptr(mean), L.ptr(control), L.ptr(logp), L.ptr(u8_out), B, L.stream_ptr(device)),
            "dxmi_var_rollout")
        extra = {"sample_u8": u8_out} if u8_out is not None else {}
        return {
            **extra,
            "sample": l_sample[T],
            "l_sample": [l_sample[i] for i in range(T + 1)],
            "logp": [logp[i] for i in range(T)],
            "logp_terminal": torch.zeros(B, device=device),
            "mean": [mean[i] for i in range(T)],
            "sigma": [sig_dev[i].repeat(B)[:, None, None, None] for i in range(T)],
            "control": [control[i] for i in range(T)],
        }

    def _sample_with_grad(self, B, device, noise):
        """Reference VARSampler.sample(enable_grad=True) (var_sampler.py:249, :411-416; train_cifar10.py:186): the rollout keeps its
        autograd graph - through the U-Net parameters, log_betas and the states.  T U-Net forwards are pending at once, so each is
        activation-checkpointed: its backward re-runs the training forward of that step before walking the tape."""
        T, shape = self.n_timesteps, tuple(self.sample_shape)
        if noise is None:
            nz = [torch.randn(B, *shape, device=device) for _ in range(T + 1)]
        else:
            nz = list(noise) if not torch.is_tensor(noise) else [noise[i] for i in range(T + 1)]
            nz = [z.to(device=device, dtype=torch.float32) for z in nz]
        net = _inner(self.net)
        x = nz[0]
        l_sample, l_mean, l_logp, l_sigma, l_control = [x], [], [], [], []
        net._checkpoint_activations = True
        try:
            with torch.enable_grad():
                for i in range(T):
                    t = torch.full((B,), i, dtype=torch.long, device=device)
                    d = self._sample_step_autograd(x, t, nz[i + 1])
                    x = d["sample"]
                    l_sample.append(x)
                    l_mean.append(d["mean"])
                    l_logp.append(d["logp"])
                    l_sigma.append(d["sigma"])
                    l_control.append(d["control"])
        finally:
            net._checkpoint_activations = False
        return {"sample": x, "l_sample": l_sample, "logp": l_logp, "logp_terminal": torch.zeros(B, device=device), "mean": l_mean,
                "sigma": l_sigma, "control": l_control}

    def _sample_step_autograd(self, x, t, noise):
        """Training form of sample_step (trainer.py:357-389 differentiates it w.r.t. the U-Net parameters and log_betas): eps
        comes from the B200 U-Net under autograd; the transition itself is a handful of elementwise torch ops so that autograd
        also reaches the learnable log_betas (reference var_sampler.py:375-408)."""
        device = x.device
        net = _inner(self.net)
        eps = self._forward_net(x, self.continuous_steps.to(device)[t])
        a = self.x_prev_multiplier.to(device)[t][:, None, None, None]
        c = (self.theta_multiplier.to(device)[t] * self.adhoc_scale1)[:, None, None, None]
        control = c * eps
        pred_mean = x * a + control
        if self.trainable_beta == "fix_last":
            log_betas_all = torch.cat([net.log_betas[:-1], net.std[-1].log().unsqueeze(0)])
            sigma = torch.exp(log_betas_all[t])
        elif self.trainable_beta:
            sigma = torch.exp(net.log_betas[t])
        else:
            sigma = self.std.to(device)[t].float()
        sigma = sigma[:, None, None, None]
        z = torch.randn_like(x) if noise is None else noise.to(device=device, dtype=torch.float32)
        xn = pred_mean + sigma * z
        logp = torch.distributions.Normal(pred_mean, sigma).log_prob(xn.detach().clone()).mean(-1).mean(-1).mean(-1)
        return {"sample": xn, "logp": logp, "logp_terminal": torch.zeros(len(x), device=device), "mean": pred_mean, "sigma": sigma,
                "entropy": torch.log(sigma), "control": control}

    def sample_step(self, x, t, y=None, noise=None):
        """Reference VARSampler.sample_step (:357-408): per-sample integer step index t."""
        device = x.device
        t = process_single_t(x, t).to(device)
        B = x.shape[0]
        x = x.detach().contiguous().float()
        net = _inner(self.net)
        if torch.is_grad_enabled() and net.training:
            return self._sample_step_autograd(x, t, noise)
        eps = self._forward_net(x, self.continuous_steps.to(device)[t])
        sig = self._sigmas().to(device)[t].float().contiguous()
        a = self.x_prev_multiplier.to(device)[t].contiguous()
        c = (self.theta_multiplier.to(device)[t] * self.adhoc_scale1).contiguous()
        z = torch.randn_like(x) if noise is None else noise.to(device=device, dtype=torch.float32).contiguous()
        xn, mean, control = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        logp = torch.empty(B, device=device)
        chw = x[0].numel()
        L.check(L.lib().dxmi_var_step(L.ptr(x), L.ptr(eps), L.ptr(z), L.ptr(a), L.ptr(c), L.ptr(sig), L.ptr(xn),
                                      L.ptr(mean), L.ptr(control), L.ptr(logp), B, chw, L.stream_ptr(x)), "dxmi_var_step")
        sigma = sig[:, None, None, None]
        return {"sample": xn, "logp": logp, "logp_terminal": torch.zeros(B, device=device), "mean": mean, "sigma": sigma,
                "entropy": torch.log(sigma), "control": control}
