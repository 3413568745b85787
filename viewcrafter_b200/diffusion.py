"""Minimal host-side model wrapper with the attribute/method surface ``DDIMSampler`` and
``utils/diffusion_utils.image_guided_synthesis`` read from the reference ``VIPLatentDiffusion``
(lvdm/models/ddpm3d.py): schedule buffers (:123-150), ``scale_arr`` (:522-527), ``apply_model`` (:723-738) with the
hybrid conditioning of ``DiffusionWrapper.forward`` (:1437-1443), ``predict_*_from_z_and_v`` (:239-251) and
``decode_first_stage`` / ``decode_core`` (:646-671).  Sub-module names (``model.diffusion_model``,
``first_stage_model``) keep the reference checkpoint prefixes.  Used by bench.py / tests / smoke when the reference
package is not importable; inside the reference repo the reference's own wrapper is used unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import schedule
from .autoencoder import AutoencoderKL
from .distributions import DiagonalGaussianDistribution
from .unet import UNetModel


class DiffusionWrapper(nn.Module):
    def __init__(self, unet: UNetModel, conditioning_key="hybrid"):
        super().__init__()
        self.diffusion_model = unet
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None, **kwargs):
        if self.conditioning_key != "hybrid":
            raise NotImplementedError("only the 'hybrid' conditioning of ViewCrafter is implemented")
        xc = torch.cat([x] + c_concat, dim=1)
        # a single entry is passed through as the SAME tensor object: the U-Net keys its cross-attention K/V cache on it
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(xc, t, context=cc, **kwargs)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config: dict, first_stage_config: dict = None, timesteps=1000, linear_start=0.00085,
                 linear_end=0.012, rescale_betas_zero_snr=True, parameterization="v", scale_factor=0.18215,
                 use_dynamic_rescale=True, base_scale=0.3, turning_step=400, perframe_ae=True, conditioning_key="hybrid",
                 decode_batch: int = 0):
        super().__init__()
        self.parameterization = parameterization
        self.scale_factor = scale_factor
        self.use_dynamic_rescale = use_dynamic_rescale
        self.perframe_ae = perframe_ae
        self.decode_batch = decode_batch          # 0 = follow perframe_ae; n>0 = decode n frames per call
        self.num_timesteps = int(timesteps)
        for k, v in schedule.model_buffers(timesteps, linear_start, linear_end, rescale_betas_zero_snr, base_scale,
                                           turning_step, use_dynamic_rescale).items():
            self.register_buffer(k, v)
        self.model = DiffusionWrapper(UNetModel(**unet_config), conditioning_key)
        self.first_stage_model = AutoencoderKL(**first_stage_config) if first_stage_config else None

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond, **kwargs):
        if not isinstance(cond, dict):
            cond = {"c_crossattn": cond if isinstance(cond, list) else [cond]}
        out = self.model(x_noisy, t, **cond, **kwargs)
        return out[0] if isinstance(out, tuple) else out

    @staticmethod
    def _gather(a, t, x):
        return a.gather(-1, t).reshape(t.shape[0], *((1,) * (x.dim() - 1)))

    def q_sample(self, x_start, t, noise=None):
        """Forward diffusion to step t (ddpm3d.py:305-308); the sampler's mask / x0 blending calls it (ddim.py:182)."""
        noise = torch.randn_like(x_start) if noise is None else noise
        return self._gather(self.sqrt_alphas_cumprod, t, x_start) * x_start + self._gather(self.sqrt_one_minus_alphas_cumprod, t, x_start) * noise

    def predict_start_from_z_and_v(self, x_t, t, v):
        return self._gather(self.sqrt_alphas_cumprod, t, x_t) * x_t - self._gather(self.sqrt_one_minus_alphas_cumprod, t, x_t) * v

    def predict_eps_from_z_and_v(self, x_t, t, v):
        return self._gather(self.sqrt_alphas_cumprod, t, x_t) * v + self._gather(self.sqrt_one_minus_alphas_cumprod, t, x_t) * x_t

    def get_first_stage_encoding(self, encoder_posterior, noise=None):
        """ddpm3d.py:611-618."""
        if isinstance(encoder_posterior, DiagonalGaussianDistribution):
            z = encoder_posterior.sample(noise=noise)
        elif isinstance(encoder_posterior, torch.Tensor):
            z = encoder_posterior
        else:
            raise NotImplementedError(f"encoder_posterior of type '{type(encoder_posterior)}' not yet implemented")
        return self.scale_factor * z

    @torch.no_grad()
    def encode_first_stage(self, x):
        """[B,3,T,H,W] (or [N,3,H,W]) -> scaled latents, one posterior sample per encode call in the reference's order
        (per frame when perframe_ae), ddpm3d.py:620-644."""
        reshape_back = x.dim() == 5
        if reshape_back:
            b, c, t, h, w = x.shape
            x = x.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        if not self.perframe_ae:
            results = self.get_first_stage_encoding(self.first_stage_model.encode(x)).detach()
        else:
            results = torch.cat([self.get_first_stage_encoding(self.first_stage_model.encode(x[i:i + 1])).detach()
                                 for i in range(x.shape[0])], dim=0)
        if reshape_back:
            results = results.reshape(b, t, *results.shape[1:]).permute(0, 2, 1, 3, 4)
        return results

    @torch.no_grad()
    def decode_core(self, z, **kwargs):
        five_d = z.dim() == 5
        if five_d:
            b, c, t, h, w = z.shape
            z = z.permute(0, 2, 1, 3, 4).reshape(b * t, c, h, w)
        step = self.decode_batch if self.decode_batch > 0 else (1 if self.perframe_ae else z.shape[0])
        outs = [self.first_stage_model.decode(1. / self.scale_factor * z[i:i + step], **kwargs) for i in range(0, z.shape[0], step)]
        r = torch.cat(outs, dim=0)
        if five_d:
            r = r.reshape(b, t, *r.shape[1:]).permute(0, 2, 1, 3, 4)
        return r

    @torch.no_grad()
    def decode_first_stage(self, z, **kwargs):
        return self.decode_core(z, **kwargs)
