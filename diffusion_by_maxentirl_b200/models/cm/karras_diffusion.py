"""Drop-in for the four hot functions of the reference's models/cm/karras_diffusion.py: `KarrasDenoiser.get_scalings`
(:64-68), `KarrasDenoiser.denoise` (:336-351), `get_sigmas_karras` (:423-429) and `get_ancestral_step` (:437-444).
The consistency-model losses and the Heun / DPM / multistep samplers of that file are not on the DxMI path."""
import torch as th

from diffusion_by_maxentirl_b200.schedule import karras_sigmas

from .nn import append_dims


def get_sigmas_karras(n, sigma_min, sigma_max, rho=7.0, device="cpu"):
    """Karras et al. (2022) noise schedule with a trailing zero."""
    return karras_sigmas(n, sigma_min, sigma_max, rho).to(device)


def get_ancestral_step(sigma_from, sigma_to):
    """sigma_down / sigma_up of one ancestral step."""
    sigma_up = (sigma_to**2 * (sigma_from**2 - sigma_to**2) / sigma_from**2) ** 0.5
    sigma_down = (sigma_to**2 - sigma_up**2) ** 0.5
    return sigma_down, sigma_up


class KarrasDenoiser:
    def __init__(self, sigma_data=0.5, sigma_max=80.0, sigma_min=0.002, rho=7.0, weight_schedule="karras",
                 distillation=False, loss_norm="l2"):
        if distillation:
            raise NotImplementedError("distillation=True (consistency boundary condition) is not used by DxMI configs")
        self.sigma_data = sigma_data
        self.sigma_max = sigma_max
        self.sigma_min = sigma_min
        self.weight_schedule = weight_schedule
        self.distillation = distillation
        self.loss_norm = loss_norm
        self.rho = rho

    def get_scalings(self, sigma):
        c_skip = self.sigma_data**2 / (sigma**2 + self.sigma_data**2)
        c_out = sigma * self.sigma_data / (sigma**2 + self.sigma_data**2) ** 0.5
        c_in = 1 / (sigma**2 + self.sigma_data**2) ** 0.5
        return c_skip, c_out, c_in

    def denoise(self, model, x_t, sigmas, **model_kwargs):
        """Returns (model_output, denoised).  The c_in scaling is folded into the U-Net's input load (`x_scale`)."""
        c_skip, c_out, c_in = self.get_scalings(sigmas)
        rescaled_t = 1000 * 0.25 * th.log(sigmas + 1e-44)
        model_output = model(x_t, rescaled_t, x_scale=c_in, **model_kwargs)
        denoised = append_dims(c_out, x_t.ndim) * model_output + append_dims(c_skip, x_t.ndim) * x_t
        return model_output, denoised
