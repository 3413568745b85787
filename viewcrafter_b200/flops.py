"""Analytical work model of the U-Net forward (2 FLOP per MAC), used for the roofline figures and to scale the
bounded CPU-baseline sample to the headline workload.  Walks the parameter tree of viewcrafter_b200.UNetModel."""
from __future__ import annotations

from .unet import UNetModel, _Down, _Res, _Transformer, _Up
import torch.nn as nn


def unet_forward_flops(m: UNetModel, T: int, H: int, W: int, ctx_len: int = 333, B: int = 1, replicate_context: bool = True) -> dict:
    """Returns dict(conv=, linear=, attention=, total=) in FLOPs.  replicate_context=True counts the cross-attention K/V
    projections on T-times replicated context rows like the reference executes them (openaimodel3d.py:562)."""
    acc = dict(conv=0.0, linear=0.0, attention=0.0)

    def tf(mod: _Transformer, h, w):
        rows = B * T * h * w
        C, inner, heads = mod.channels, mod.heads * 64, mod.heads
        acc["linear"] += 2 * rows * C * inner * 2                      # proj_in + proj_out
        for _ in mod.transformer_blocks:
            acc["linear"] += 2 * rows * inner * inner * 4              # attn1 q,k,v,out
            acc["linear"] += 2 * rows * inner * inner * 2              # attn2 q,out
            acc["linear"] += 2 * rows * inner * (8 * inner) + 2 * rows * (4 * inner) * inner   # GEGLU FF
            if mod.kind == "S":
                n = h * w
                acc["attention"] += 4 * B * T * heads * n * n * 64
                acc["attention"] += 4 * B * T * heads * n * ctx_len * 64
                ctx_rows = B * (T if replicate_context else 1) * ctx_len
                acc["linear"] += 2 * ctx_rows * 1024 * inner * 2
            else:
                acc["linear"] += 2 * rows * inner * inner * 2          # attn2 k,v (self-attention again)
                acc["attention"] += 2 * 4 * B * h * w * heads * T * T * 64

    def res(mod: _Res, h, w):
        rows = B * T * h * w
        acc["conv"] += 2 * rows * 9 * mod.cin * mod.cout + 2 * rows * 9 * mod.cout * mod.cout
        if isinstance(mod.skip_connection, nn.Conv2d):
            acc["conv"] += 2 * rows * mod.cin * mod.cout
        acc["linear"] += 2 * B * T * 4 * m.model_channels * mod.cout
        if hasattr(mod, "temopral_conv"):
            acc["conv"] += 4 * 2 * rows * 3 * mod.cout * mod.cout

    def stage(st, h, w):
        for mod in st:
            if isinstance(mod, _Res):
                res(mod, h, w)
            elif isinstance(mod, _Transformer):
                tf(mod, h, w)
            elif isinstance(mod, _Down):
                h, w = (h + 1) // 2, (w + 1) // 2
                acc["conv"] += 2 * B * T * h * w * 9 * mod.op.in_channels * mod.op.out_channels
            elif isinstance(mod, _Up):
                h, w = 2 * h, 2 * w
                acc["conv"] += 2 * B * T * h * w * 9 * mod.conv.in_channels * mod.conv.out_channels
            elif isinstance(mod, nn.Conv2d):
                acc["conv"] += 2 * B * T * h * w * 9 * mod.in_channels * mod.out_channels
        return h, w

    h, w = H, W
    for i, st in enumerate(m.input_blocks):
        h, w = stage(st, h, w)
        if i == 0 and m.addition_attention:
            stage(m.init_attn, h, w)
    h, w = stage(m.middle_block, h, w)
    for st in m.output_blocks:
        h, w = stage(st, h, w)
    acc["conv"] += 2 * B * T * h * w * 9 * m.model_channels * m.out_channels
    acc["linear"] += 2 * B * (m.model_channels * 4 * m.model_channels + (4 * m.model_channels) ** 2) * (2 if m.fs_condition else 1)
    acc["total"] = acc["conv"] + acc["linear"] + acc["attention"]
    return acc
