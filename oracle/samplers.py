"""Oracle (CPU): the two DxMI samplers - schedule pre-computation and the T-step rollouts.

TEST INFRASTRUCTURE - see oracle/__init__.py.  Restates models/DxMI/var_sampler.py and
models/DxMI/openai_diffusion.py (+ the four hot functions of models/cm/karras_diffusion.py).
Noise is always host-supplied: `noise[0]` is x_0 (before the sigma_max scaling for EDM: the caller passes the
already-scaled x_0), `noise[1 + i]` is the z drawn in step i.
"""
import math

import numpy as np
import torch

# ------------------------------------------------------------------------------------------------ VARSampler schedule

BETA_0, BETA_T, T_TRAIN = 0.0001, 0.02, 1000  # var_sampler.py:13-17


def _bisect(f, lo, hi, target, eps):
    """var_sampler.py:47-70: f is decreasing; stop when target <= f(x) <= (1 +- eps) * target."""
    sign = -1.0 if target < 0 else 1.0
    x = None
    for _ in range(1000):
        x = (lo + hi) / 2
        fx = f(x)
        if fx < target:
            hi = x
        elif fx > (1 + sign * eps) * target:
            lo = x
        else:
            break
    return x


def var_eta(T_user):
    """var_sampler.py:73-97 with schedule='quadratic': eta_i = beta_0 (1 + i x)^2, x chosen so that
    prod(1 - eta) matches alpha_bar_T of the 1000-step linear schedule (float64 throughout)."""
    target = float(np.prod(1 - np.linspace(BETA_0, BETA_T, T_TRAIN)))

    def g(x):
        return np.array([BETA_0 * (1 + i * x) ** 2 for i in range(T_user)])

    x = _bisect(lambda v: float(np.prod(1 - g(v))), 0.0, 0.95 / math.sqrt(BETA_0) / T_user, target, 1e-4)
    return g(x)


def _log_gamma_stirling(x):
    """var_sampler.py:100-103."""
    y = x - 1
    return math.log(2 * math.pi * y) / 2 + y * (math.log(y) - 1) + math.log(1 + 1 / (12 * y))


def _log_alpha_bar_continuous(t, beta0_f32, betaT_f32):
    """var_sampler.py:106-111 under numpy<2 value-based promotion: the float32 difference
    beta_T - beta_0 stays float32, every mixed scalar operation after it is float64."""
    delta = float(np.float32(betaT_f32) - np.float32(beta0_f32)) / (T_TRAIN - 1)
    c = (1.0 - float(beta0_f32)) / delta
    t1 = t + 1
    return t1 * math.log(delta) + _log_gamma_stirling(c + 1) - _log_gamma_stirling(c - t1 + 1)


def var_schedule(T_user, kappa=1.0):
    """VARSampler.init_schedule + VAR_get_params (var_sampler.py:326-355, :146-186, :115-143).
    Returns a dict of the sampler buffers (float32 torch tensors, same arithmetic order as the reference)."""
    eta = var_eta(T_user)
    beta = torch.linspace(BETA_0, BETA_T, T_TRAIN)  # float32, calc_diffusion_hyperparams :33-39
    alpha_bar = 1 - beta
    for t in range(1, T_TRAIN):
        alpha_bar[t] *= alpha_bar[t - 1]
    gamma_bar = 1 - torch.from_numpy(eta).to(torch.float32)
    for t in range(1, T_user):
        gamma_bar[t] *= gamma_bar[t - 1]
    assert gamma_bar[0] <= alpha_bar[0] and gamma_bar[-1] >= alpha_bar[-1]

    b0, bT = beta[0].numpy(), beta[-1].numpy()
    steps = []
    for t in range(T_user - 1, -1, -1):  # _precompute_VAR_steps :131-143
        tau = None
        for i in range(T_TRAIN - 1):
            if alpha_bar[i] >= gamma_bar[t] > alpha_bar[i + 1]:
                target = float(np.log(gamma_bar[t].numpy()))  # float32 log, then promoted
                tau = _bisect(lambda v: _log_alpha_bar_continuous(v, b0, bT), i - 0.01, i + 1.01, target, 1e-8)
                break
        if tau is None:
            tau = T_TRAIN - 1
        steps.append(tau)
    continuous_steps = torch.tensor(steps)  # float32

    a = torch.zeros(T_user)
    c = torch.zeros(T_user)
    std = torch.zeros(T_user)
    for i, tau in enumerate(steps):  # VAR_get_params :165-183
        j = T_user - 1 - i
        if i == T_user - 1:
            assert abs(tau) < 0.1
            alpha_next, sigma = torch.tensor(1.0), torch.tensor(0.0)
        else:
            alpha_next = gamma_bar[j - 1]
            sigma = kappa * torch.sqrt((1 - alpha_next) / (1 - gamma_bar[j]) * (1 - gamma_bar[j] / alpha_next))
        a[i] = torch.sqrt(alpha_next / gamma_bar[j])
        c[i] = torch.sqrt(1 - alpha_next - sigma**2) - torch.sqrt(1 - gamma_bar[j]) * torch.sqrt(alpha_next / gamma_bar[j])
        std[i] = 0.001 if i == T_user - 1 else sigma
    return {
        "user_defined_eta": eta,
        "continuous_steps": continuous_steps,
        "Gamma_bar": gamma_bar,
        "x_prev_multiplier": a,
        "theta_multiplier": c,
        "std": std,
        "log_betas_init": torch.log(std),  # init_schedule :341-355 (adhoc_scale2 = 1)
    }


def var_sigmas(log_betas, std, trainable_beta="fix_last"):
    """The noise scale actually used per step (var_sampler.py:268-283)."""
    if trainable_beta == "fix_last":
        return torch.exp(torch.cat([log_betas[:-1], std[-1].log().unsqueeze(0)]))
    if trainable_beta:
        return torch.exp(log_betas)
    return std.clone()


def var_rollout(net, sched, log_betas, noise, trainable_beta="fix_last", adhoc_scale1=1.0):
    """VAR_sampling (var_sampler.py:204-297) / VARSampler.sample (:411-428) on host-supplied noise.
    net(x, t) -> eps.  Returns the reference's d_sample dict."""
    T = len(sched["continuous_steps"])
    sig = var_sigmas(log_betas, sched["std"], trainable_beta)
    x = noise[0].clone()
    B = x.shape[0]
    out = {"l_sample": [x.clone()], "logp": [], "control": [], "mean": [], "sigma": [], "eps": []}
    for i in range(T):
        tau = sched["continuous_steps"][i]
        eps = net(x, tau * torch.ones(B))
        x = x * sched["x_prev_multiplier"][i]
        control = sched["theta_multiplier"][i] * eps * adhoc_scale1
        mean = x + control
        s = sig[i]
        x = x + (control + s * noise[i + 1])
        logp = (-((x - mean) ** 2) / (2 * s * s) - torch.log(s) - math.log(math.sqrt(2 * math.pi))).mean(-1).mean(-1).mean(-1)
        out["l_sample"].append(x.clone())
        out["logp"].append(logp)
        out["control"].append(control)
        out["mean"].append(mean)
        out["sigma"].append(s.repeat(B)[:, None, None, None])
        out["eps"].append(eps)
    out["sample"] = out["l_sample"][-1]
    out["logp_terminal"] = torch.zeros(B)
    return out


def var_sample_step(net, sched, log_betas, x, t, z, trainable_beta="fix_last", adhoc_scale1=1.0):
    """VARSampler.sample_step (var_sampler.py:357-408) for a per-sample integer step index t [B]."""
    sig = var_sigmas(log_betas, sched["std"], trainable_beta)
    eps = net(x, sched["continuous_steps"][t])
    a = sched["x_prev_multiplier"][t][:, None, None, None]
    c = sched["theta_multiplier"][t][:, None, None, None]
    s = sig[t][:, None, None, None]
    control = c * eps * adhoc_scale1
    mean = x * a + control
    xn = mean + s * z
    logp = (-((xn - mean) ** 2) / (2 * s * s) - torch.log(s) - math.log(math.sqrt(2 * math.pi))).mean(-1).mean(-1).mean(-1)
    return {"sample": xn, "logp": logp, "logp_terminal": torch.zeros(len(x)), "mean": mean, "sigma": s,
            "entropy": torch.log(s), "control": control}


# ------------------------------------------------------------------------------------------------ EDM / Karras

SIGMA_DATA = 0.5


def karras_sigmas(n, sigma_min, sigma_max, rho):
    """models/cm/karras_diffusion.py:423-429 (with the appended zero)."""
    ramp = torch.linspace(0, 1, n)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    sig = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([sig, sig.new_zeros([1])])


def edm_schedule(n_timesteps, sigma_min=0.002, sigma_max=80.0, rho=7.0, stochastic_last=False):
    """OpenAIDiffusion.__init__ (models/DxMI/openai_diffusion.py:29-56)."""
    if stochastic_last:
        sigmas = karras_sigmas(n_timesteps + 1, sigma_min, sigma_max, rho)[:-1]
    else:
        sigmas = karras_sigmas(n_timesteps, sigma_min, sigma_max, rho)
    s_from, s_to = sigmas[:-1], sigmas[1:]
    sigma_up = (s_to**2 * (s_from**2 - s_to**2) / s_from**2) ** 0.5
    sigma_down = (s_to**2 - sigma_up**2) ** 0.5
    return {"sigmas": sigmas, "sigma_up": sigma_up, "sigma_down": sigma_down,
            "log_betas_init": torch.log(sigma_up.clamp(1e-3))}


def edm_scalings(sigma):
    """karras_diffusion.py:64-68."""
    c_skip = SIGMA_DATA**2 / (sigma**2 + SIGMA_DATA**2)
    c_out = sigma * SIGMA_DATA / (sigma**2 + SIGMA_DATA**2) ** 0.5
    c_in = 1 / (sigma**2 + SIGMA_DATA**2) ** 0.5
    return c_skip, c_out, c_in


def edm_noise_sigma(sched, log_betas, i, n_timesteps, trainable_beta="fix_last"):
    """openai_diffusion.py:79-92 for a scalar step index."""
    if not trainable_beta:
        return sched["sigma_up"][i]
    s = torch.exp(log_betas[i])
    if trainable_beta == "fix_last" and i == n_timesteps - 1:
        s = sched["sigma_up"][i]
    elif trainable_beta == "fix_last3" and not (i < n_timesteps - 3):
        s = sched["sigma_up"][i]
    return s


def edm_rollout(net, sched, log_betas, noise, y=None, trainable_beta="fix_last"):
    """OpenAIDiffusion.sample (openai_diffusion.py:101-127) with KarrasDenoiser.denoise (karras_diffusion.py:336-351).
    net(x_in, rescaled_t, y) -> F.  noise[0] = x_0 = sigma_max * z_0."""
    T = len(sched["sigma_up"])
    x = noise[0]
    B = x.shape[0]
    out = {"l_sample": [x], "mean": [], "sigma": [], "F": [], "y": y}
    for i in range(T):
        sigma = sched["sigmas"][i]
        c_skip, c_out, c_in = edm_scalings(sigma)
        rescaled_t = 1000 * 0.25 * torch.log(sigma * torch.ones(B) + 1e-44)
        Fx = net(c_in * x, rescaled_t, y)
        denoised = c_out * Fx + c_skip * x
        d = (x - denoised) / sigma
        mu = x + d * (sched["sigma_down"][i] - sigma)
        s_noise = edm_noise_sigma(sched, log_betas, i, T, trainable_beta)
        x = mu + noise[i + 1] * s_noise
        out["l_sample"].append(x)
        out["mean"].append(mu)
        out["sigma"].append((s_noise * torch.ones(B)).clamp(1e-4, None))
        out["F"].append(Fx)
    out["sample"] = x
    return out
