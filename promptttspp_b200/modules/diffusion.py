"""DDPM sampler buffers (reference: promptttspp/modules/diffusion.py:71-163).

The twelve [K_step] schedule buffers are part of the checkpoint; they are
recomputed here in float64 exactly as the reference does and the sampler
(csrc/acoustic.cu::diffusion_sample) reads whatever ``load_state_dict`` left
in them.
"""
import numpy as np
import torch
from torch import nn


def _linear_betas(timesteps, min_beta=1e-4, max_beta=0.06):
    return np.linspace(min_beta, max_beta, timesteps)


def _cosine_betas(timesteps, s=0.008):
    steps = timesteps + 1
    x = np.linspace(0, steps, steps)
    ac = np.cos(((x / steps) + s) / (1 + s) * np.pi * 0.5) ** 2
    ac = ac / ac[0]
    return np.clip(1 - (ac[1:] / ac[:-1]), a_min=0, a_max=0.999)


class GaussianDiffusion(nn.Module):
    def __init__(self, in_dim, out_dim, denoise_fn, encoder=None, K_step=100, betas=None,
                 schedule_type="linear", scheduler_params=None, norm_scale=None,
                 a_min=0, a_max=20, pndm_speedup=None):
        super().__init__()
        if pndm_speedup:
            raise NotImplementedError("pndm_speedup is not implemented yet")
        if encoder is not None:
            raise NotImplementedError("GaussianDiffusion(encoder=...) is not used by the shipped configs")
        assert out_dim == denoise_fn.in_dim, "denoise_fn input dim must match out_dim"
        self.in_dim, self.out_dim = in_dim, out_dim
        self.denoise_fn = denoise_fn
        self.K_step = K_step
        self.norm_scale, self.a_min, self.a_max = norm_scale, a_min, a_max
        if betas is not None:
            betas = betas.detach().cpu().numpy() if isinstance(betas, torch.Tensor) else betas
        elif schedule_type == "linear":
            betas = _linear_betas(K_step, **(scheduler_params or {"max_beta": 0.06}))
        elif schedule_type == "cosine":
            betas = _cosine_betas(K_step, **(scheduler_params or {"s": 0.008}))
        else:
            raise ValueError(schedule_type)
        alphas = 1.0 - betas
        ac = np.cumprod(alphas, axis=0)
        ac_prev = np.append(1.0, ac[:-1])
        post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
        table = {
            "betas": betas,
            "alphas_cumprod": ac,
            "alphas_cumprod_prev": ac_prev,
            "sqrt_alphas_cumprod": np.sqrt(ac),
            "sqrt_one_minus_alphas_cumprod": np.sqrt(1.0 - ac),
            "log_one_minus_alphas_cumprod": np.log(1.0 - ac),
            "sqrt_recip_alphas_cumprod": np.sqrt(1.0 / ac),
            "sqrt_recipm1_alphas_cumprod": np.sqrt(1.0 / ac - 1),
            "posterior_variance": post_var,
            "posterior_log_variance_clipped": np.log(np.maximum(post_var, 1e-20)),
            "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
            "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        }
        for name, val in table.items():
            self.register_buffer(name, torch.tensor(val, dtype=torch.float32))
