"""Drop-in ``DDIMSampler`` with three-way classifier-free guidance (reference: lvdm/models/samplers/ddim_multiplecond.py),
the sampler ``image_guided_synthesis(..., multiple_cond_cfg=True)`` selects (utils/diffusion_utils.py:9,119).

Differences from ``viewcrafter_b200.ddim.DDIMSampler`` -- exactly the reference's:
  * three ``apply_model`` calls per step: cond, uncond and ``unconditional_conditioning_img_nonetext`` (image kept, text
    dropped), combined as ``u + cfg_img (v_img - u) + s (v_cond - v_img)`` (ddim_multiplecond.py:227-233);
    ``cfg_img`` defaults to the text scale;
  * ``ddim_scale_arr_prev[0] = ddim_scale_arr[0]`` (ddim_multiplecond.py:33; ddim.py:33-35 fixed this one only), so the last
    step's dynamic rescale differs between the two samplers.
The combine, guidance rescale, v->(eps, x0), dynamic rescale and x_{t-1} are one fused CUDA update (vc_ddim_update3).
"""
from __future__ import annotations

import torch

from . import ops
from .ddim import DDIMSampler as _TwoWaySampler


class DDIMSampler(_TwoWaySampler):
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0., verbose=True):
        super().make_schedule(ddim_num_steps, ddim_discretize, ddim_eta, verbose)
        if self.use_dynamic_rescale:
            self.ddim_scale_arr_prev = torch.cat([self.ddim_scale_arr[0:1], self.ddim_scale_arr[:-1]])

    @torch.no_grad()
    def p_sample_ddim(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1., noise_dropout=0., score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1., unconditional_conditioning=None, uc_type=None, cfg_img=None,
                      mask=None, x0=None, guidance_rescale=0.0, _step=None, **kwargs):
        self._check_step_options(use_original_steps, quantize_denoised, score_corrector)
        if getattr(self.model, "_cfg", None) is not None:
            raise NotImplementedError("viewcrafter_b200.DDIMSampler(multicond): the 2-way CFG rank split does not cover three branches")
        if cfg_img is None:
            cfg_img = unconditional_guidance_scale
        uc_img = kwargs['unconditional_conditioning_img_nonetext']           # KeyError like ddim_multiplecond.py:224
        step = int(t[0].item()) if _step is None else _step
        v_u = v_i = None
        if unconditional_conditioning is None or unconditional_guidance_scale == 1.:
            v_c = self.model.apply_model(x, t, c, **kwargs)
        else:
            if uc_img is None:
                raise ValueError("three-way CFG needs unconditional_conditioning_img_nonetext (image_guided_synthesis only builds it "
                                 "when cfg_img != 1.0, utils/diffusion_utils.py:157-163)")
            v_c, v_u = self._apply_both(x, t, c, unconditional_conditioning, kwargs)
            v_i = self.model.apply_model(x, t, uc_img, **kwargs)
        sc = self.step_scalars(index, step)
        sc["cfg_scale"], sc["guidance_rescale"] = float(unconditional_guidance_scale), float(guidance_rescale)
        noise = self._step_noise(x, repeat_noise, temperature, noise_dropout)
        return self._fused_update(x, v_c, v_u, noise, sc, v_uncond_img=v_i, cfg_img=float(cfg_img))
