"""Mirror of lvdm/distributions.py:24-87 (DiagonalGaussianDistribution): same attributes, same RNG behaviour
(`sample()` draws `torch.randn(mean.shape)` on the CPU generator and moves it to the parameters' device).

Inside the reference repo the caller type-checks the posterior -- ``LatentDiffusion.get_first_stage_encoding`` does
``isinstance(encoder_posterior, lvdm.distributions.DiagonalGaussianDistribution)`` (ddpm3d.py:611-618) -- so
``posterior_class()`` hands ``AutoencoderKL.encode`` a subclass of BOTH this mirror and the reference's class whenever
``lvdm.distributions`` has been imported in the process (it is, by the time ddpm3d.py runs)."""
from __future__ import annotations

import sys

import torch


class DiagonalGaussianDistribution(object):
    """Posterior q(z|x) = N(mean, diag(exp(logvar))) over the VAE moments ``[N, 2*embed, h, w]``.

    Inference surface only -- what ``get_first_stage_encoding`` / ``encode_first_stage`` touch (ddpm3d.py:611-644):
    ``parameters``, ``mean``, ``logvar`` (clamped to [-30, 20] like distributions.py:28), ``std``, ``var``,
    ``deterministic``, ``sample(noise=None)`` and ``mode()``.  The training-time ``kl`` / ``nll`` terms are not on the
    ViewCrafter inference path; inside the reference repo they are inherited from the reference class (posterior_class)."""

    LOGVAR_RANGE = (-30.0, 20.0)

    def __init__(self, parameters, deterministic=False):
        half = parameters.shape[1] // 2
        self.parameters = parameters
        self.deterministic = bool(deterministic)
        self.mean = parameters[:, :half]
        self.logvar = parameters[:, half:].clamp(*self.LOGVAR_RANGE)
        if self.deterministic:
            self.std = self.var = torch.zeros_like(self.mean)
        else:
            self.var = self.logvar.exp()
            self.std = (0.5 * self.logvar).exp()

    def sample(self, noise=None):
        # the reference draws on the CPU generator with the shape of the mean and only then moves the draw to the device
        # (distributions.py:35-36); the order of draws is part of the drop-in contract (tests replay it)
        eps = torch.randn(self.mean.shape) if noise is None else noise
        return self.mean + eps.to(device=self.parameters.device) * self.std

    def mode(self):
        return self.mean


_subclass_cache = {}


def posterior_class():
    """The class AutoencoderKL.encode instantiates: this mirror, made a subclass of the reference's
    DiagonalGaussianDistribution when that module is loaded (so the reference's isinstance checks accept it)."""
    ref = sys.modules.get("lvdm.distributions")
    base = getattr(ref, "DiagonalGaussianDistribution", None) if ref is not None else None
    if base is None or base is DiagonalGaussianDistribution:
        return DiagonalGaussianDistribution
    cls = _subclass_cache.get(base)
    if cls is None:
        cls = type("DiagonalGaussianDistribution", (DiagonalGaussianDistribution, base), {})
        _subclass_cache[base] = cls
    return cls
