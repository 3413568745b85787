"""Host-side (numpy, float64 -> float32) diffusion schedule tables.  Bit-exactness against the reference is
checked by tests/test_schedule_cpu.py using known answers generated from the reference's own functions.

Mirrors: make_beta_schedule / rescale_zero_terminal_snr / make_ddim_timesteps / make_ddim_sampling_parameters
(lvdm/models/utils_diffusion.py:31-91,112-144), DDPM.register_schedule (lvdm/models/ddpm3d.py:123-150),
LatentDiffusion scale_arr (ddpm3d.py:522-527) and DDIMSampler.make_schedule (lvdm/models/samplers/ddim.py:24-59).
"""
from __future__ import annotations

import numpy as np
import torch


def linear_betas(n: int, start: float, end: float) -> np.ndarray:
    return (torch.linspace(start ** 0.5, end ** 0.5, n, dtype=torch.float64, device="cpu") ** 2).numpy()


def zero_terminal_snr(betas: np.ndarray) -> np.ndarray:
    root = np.sqrt(np.cumprod(1.0 - betas, axis=0))
    r0, rT = root[0].copy(), root[-1].copy()
    root = (root - rT) * (r0 / (r0 - rT))
    bar = root ** 2
    return 1 - np.concatenate([bar[0:1], bar[1:] / bar[:-1]])


def model_buffers(timesteps=1000, linear_start=0.00085, linear_end=0.012, zero_snr=True, base_scale=0.3,
                  turning_step=400, dynamic_rescale=True) -> dict:
    betas = linear_betas(timesteps, linear_start, linear_end)
    if zero_snr:
        betas = zero_terminal_snr(betas)
    ac = np.cumprod(1.0 - betas, axis=0)
    t32 = lambda a: torch.tensor(a, dtype=torch.float32, device="cpu")
    out = dict(betas=t32(betas), alphas_cumprod=t32(ac), alphas_cumprod_prev=t32(np.append(1.0, ac[:-1])),
               sqrt_alphas_cumprod=t32(np.sqrt(ac)), sqrt_one_minus_alphas_cumprod=t32(np.sqrt(1.0 - ac)))
    if dynamic_rescale:
        out["scale_arr"] = t32(np.concatenate((np.linspace(1.0, base_scale, turning_step), np.full(timesteps, base_scale))))
    return out


def ddim_timesteps(method: str, n_ddim: int, n_ddpm: int) -> np.ndarray:
    if method == "uniform":
        return np.asarray(list(range(0, n_ddpm, n_ddpm // n_ddim))) + 1
    if method == "uniform_trailing":
        return np.flip(np.round(np.arange(n_ddpm, 0, -(n_ddpm / n_ddim)))).astype(np.int64) - 1
    if method == "quad":
        return ((np.linspace(0, np.sqrt(n_ddpm * .8), n_ddim)) ** 2).astype(int) + 1
    raise NotImplementedError(f'There is no ddim discretization method called "{method}"')


def ddim_parameters(alphacums: torch.Tensor, ts: np.ndarray, eta: float):
    """Returns (sigmas float64 tensor, alphas float32 tensor, alphas_prev float64 ndarray): the reference's dtypes."""
    alphas = alphacums[ts]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ts[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def f32(v) -> float:
    """The value torch.full(size, v) would hold (float32 rounding of a python/numpy/tensor scalar)."""
    return float(np.float32(float(v)))
