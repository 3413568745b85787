"""Mirror of lvdm/distributions.py:24-87 (DiagonalGaussianDistribution): same attributes, same RNG behaviour
(`sample()` draws `torch.randn(mean.shape)` on the CPU generator and moves it to the parameters' device).

Inside the reference repo the caller type-checks the posterior -- ``LatentDiffusion.get_first_stage_encoding`` does
``isinstance(encoder_posterior, lvdm.distributions.DiagonalGaussianDistribution)`` (ddpm3d.py:611-618) -- so
``posterior_class()`` hands ``AutoencoderKL.encode`` a subclass of BOTH this mirror and the reference's class whenever
``lvdm.distributions`` has been imported in the process (it is, by the time ddpm3d.py runs)."""
from __future__ import annotations

import sys

import numpy as np
import torch


class DiagonalGaussianDistribution(object):
    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if self.deterministic:
            self.var = self.std = torch.zeros_like(self.mean).to(device=self.parameters.device)

    def sample(self, noise=None):
        if noise is None:
            noise = torch.randn(self.mean.shape)
        return self.mean + self.std * noise.to(device=self.parameters.device)

    def mode(self):
        return self.mean

    def kl(self, other=None):
        if self.deterministic:
            return torch.Tensor([0.])
        if other is None:
            return 0.5 * torch.sum(torch.pow(self.mean, 2) + self.var - 1.0 - self.logvar, dim=[1, 2, 3])
        return 0.5 * torch.sum(torch.pow(self.mean - other.mean, 2) / other.var + self.var / other.var - 1.0 - self.logvar + other.logvar,
                               dim=[1, 2, 3])

    def nll(self, sample, dims=[1, 2, 3]):
        if self.deterministic:
            return torch.Tensor([0.])
        logtwopi = np.log(2.0 * np.pi)
        return 0.5 * torch.sum(logtwopi + self.logvar + torch.pow(sample - self.mean, 2) / self.var, dim=dims)


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
