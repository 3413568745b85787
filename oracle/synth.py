"""Deterministic synthetic weights / inputs shared by the golden generator and the tests.

TEST INFRASTRUCTURE (see oracle/lvdm_oracle.py header).  There is no checkpoint on disk and no
network, so every parity test runs on seeded synthetic weights.  Weights are regenerated from
(name, shape, seed) instead of being stored, so golden fixtures only hold inputs and outputs.
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Iterable, Tuple

import torch


def synth_tensor(name: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """One tensor, independent of every other (seeded by crc32(name)): N(0,1/fan_in) for matrices and
    conv kernels, 1+0.1*N for norm scales, 0.05*N for biases."""
    g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
    shape = tuple(int(s) for s in shape)
    r = torch.randn(shape, generator=g)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return r / math.sqrt(fan_in)
    if name.endswith(".weight"):
        return 1.0 + 0.1 * r
    return 0.05 * r


def synth_state_dict(shapes: Iterable[Tuple[str, Tuple[int, ...]]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {n: synth_tensor(n, s, seed) for n, s in shapes}


def module_shapes(module: torch.nn.Module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]
