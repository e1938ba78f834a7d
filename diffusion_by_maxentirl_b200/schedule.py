"""Host-side, one-time schedule tables of the two DxMI samplers.

These are the reference's `VARSampler.init_schedule` / `VAR_get_params` (models/DxMI/var_sampler.py:326-355,
:146-186, :115-143, :73-97) and `OpenAIDiffusion.__init__` (models/DxMI/openai_diffusion.py:29-56,
models/cm/karras_diffusion.py:423-429).  They run once per sampler on the CPU (negligible), produce a handful of
[T] tables, and everything downstream (the CUDA rollout) only consumes the tables.

Numerics note: the reference relied on numpy<2 value-based promotion inside its bisection
(`float32 - float32` stays float32, every mixed scalar op is float64).  The tables here are computed with those
semantics spelled out explicitly, so they do not depend on the installed numpy.
"""
import math

import numpy as np
import torch

_BETA_0, _BETA_T, _T_TRAIN = 1e-4, 0.02, 1000


def _solve_decreasing(fn, lo, hi, target, tol, iters=1000):
    """Midpoint search for the reference's acceptance band  target <= fn(x) <= (1 +- tol) target."""
    upper = (1.0 + math.copysign(tol, target)) * target
    mid = 0.5 * (lo + hi)
    for _ in range(iters):
        mid = 0.5 * (lo + hi)
        v = fn(mid)
        if v < target:
            hi = mid
        elif v > upper:
            lo = mid
        else:
            return mid
    return mid


def quadratic_eta(T):
    """eta_i = beta_0 (1 + i s)^2 with s such that prod(1 - eta) ~= alpha_bar of the 1000-step linear schedule."""
    target = float(np.prod(1.0 - np.linspace(_BETA_0, _BETA_T, _T_TRAIN)))

    def eta_of(s):
        return np.array([_BETA_0 * (1 + i * s) ** 2 for i in range(T)])

    s = _solve_decreasing(lambda v: float(np.prod(1 - eta_of(v))), 0.0, 0.95 / math.sqrt(_BETA_0) / T, target, 1e-4)
    return eta_of(s)


def _stirling_lgamma(x):
    """log Gamma(x) by Stirling's series, as the reference does (var_sampler.py:100-103)."""
    y = x - 1.0
    return math.log(2 * math.pi * y) / 2 + y * (math.log(y) - 1) + math.log(1 + 1 / (12 * y))


class VarSchedule:
    """All [T] tables of the DDPM-style few-step sampler (float32 torch tensors on the CPU)."""

    def __init__(self, n_timesteps, kappa=1.0):
        T = int(n_timesteps)
        self.T = T
        self.kappa = kappa
        self.user_defined_eta = quadratic_eta(T)

        beta = torch.linspace(_BETA_0, _BETA_T, _T_TRAIN)
        # sequential fp32 running product, element by element like the reference
        ab = (1 - beta).clone()
        for t in range(1, _T_TRAIN):
            ab[t] = ab[t] * ab[t - 1]
        alpha_bar = ab
        gb = (1 - torch.from_numpy(self.user_defined_eta).to(torch.float32)).clone()
        for t in range(1, T):
            gb[t] = gb[t] * gb[t - 1]
        if not (gb[0] <= alpha_bar[0] and gb[-1] >= alpha_bar[-1]):
            raise ValueError("user-defined noise schedule falls outside the training schedule")
        self.Gamma_bar = gb
        self.alpha_bar = alpha_bar

        # continuous time tau with alpha_bar_cont(tau) = Gamma_bar[t]  (first step = noisiest)
        b0 = float(beta[0])
        delta = float((beta[-1] - beta[0]).item()) / (_T_TRAIN - 1)  # fp32 difference, then float64
        cc = (1.0 - b0) / delta
        log_delta = math.log(delta)

        def log_abar(tt):
            t1 = tt + 1
            return t1 * log_delta + _stirling_lgamma(cc + 1) - _stirling_lgamma(cc - t1 + 1)

        ab_list = alpha_bar.tolist()
        taus = []
        for t in range(T - 1, -1, -1):
            g32 = gb[t]
            g = float(g32)
            tau = None
            # bracket: alpha_bar[i] >= g > alpha_bar[i+1]
            lo_i, hi_i = 0, _T_TRAIN - 1
            if ab_list[0] >= g > ab_list[-1]:
                while hi_i - lo_i > 1:
                    m = (lo_i + hi_i) // 2
                    if ab_list[m] >= g:
                        lo_i = m
                    else:
                        hi_i = m
                i = lo_i
                target = float(np.log(g32.numpy()))  # float32 log (numpy's), promoted to float64 afterwards
                tau = _solve_decreasing(log_abar, i - 0.01, i + 1.01, target, 1e-8)
            if tau is None:
                tau = _T_TRAIN - 1
            taus.append(tau)
        self.continuous_steps = torch.tensor(taus)

        a = torch.zeros(T)
        c = torch.zeros(T)
        std = torch.zeros(T)
        one, zero = torch.tensor(1.0), torch.tensor(0.0)
        for i in range(T):
            j = T - 1 - i
            last = i == T - 1
            if last and not abs(taus[i]) < 0.1:
                raise ValueError("last continuous step must be ~0")
            an = one if last else gb[j - 1]
            sg = zero if last else kappa * torch.sqrt((1 - an) / (1 - gb[j]) * (1 - gb[j] / an))
            a[i] = torch.sqrt(an / gb[j])
            c[i] = torch.sqrt(1 - an - sg**2) - torch.sqrt(1 - gb[j]) * torch.sqrt(an / gb[j])
            std[i] = 0.001 if last else sg
        self.x_prev_multiplier = a
        self.theta_multiplier = c
        self.std = std
        self.diffusion_steps_list = self.continuous_steps.clone().to(torch.float32)


def karras_sigmas(n, sigma_min=0.002, sigma_max=80.0, rho=7.0):
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([sig, sig.new_zeros([1])])


class EdmSchedule:
    """sigmas / sigma_down / sigma_up of the ancestral EDM sampler (float32 torch tensors on the CPU)."""

    def __init__(self, n_timesteps, sigma_min=0.002, sigma_max=80.0, rho=7.0, stochastic_last=False):
        if stochastic_last:
            self.sigmas = karras_sigmas(n_timesteps + 1, sigma_min, sigma_max, rho)[:-1]
        else:
            self.sigmas = karras_sigmas(n_timesteps, sigma_min, sigma_max, rho)
        s_from, s_to = self.sigmas[:-1], self.sigmas[1:]
        self.sigma_up = (s_to**2 * (s_from**2 - s_to**2) / s_from**2) ** 0.5
        self.sigma_down = (s_to**2 - self.sigma_up**2) ** 0.5

    @staticmethod
    def scalings(sigma, sigma_data=0.5):
        """c_skip, c_out, c_in (karras_diffusion.py:64-68) for a float32 tensor of sigmas."""
        c_skip = sigma_data**2 / (sigma**2 + sigma_data**2)
        c_out = sigma * sigma_data / (sigma**2 + sigma_data**2) ** 0.5
        c_in = 1 / (sigma**2 + sigma_data**2) ** 0.5
        return c_skip, c_out, c_in
