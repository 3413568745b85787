"""TEST DOUBLE for viewcrafter_b200.ops: the same call surface implemented with plain torch on the CPU (fp32 maths,
fp16 storage).  Lets the CPU suite (-m "not gpu") exercise the HOST logic of the drop-in classes -- block wiring,
weight packing, row layouts, K-split concat, per-batch loops -- against the oracle without a GPU.  It lives under
tests/ and is never imported by the product; the product path has no fallback and fails loudly without CUDA."""
from __future__ import annotations

import math
import sys

import numpy as np
import torch
import torch.nn.functional as F

from viewcrafter_b200 import ops as real

pack_conv3x3 = real.pack_conv3x3
pack_conv_temporal = real.pack_conv_temporal
pack_linear = real.pack_linear
pack_upconv3x3 = real.pack_upconv3x3


def require_cuda(device, who):
    return None


def _geglu_tile(N):
    return 256 if N % 256 == 0 else 128


def pack_geglu(w, b):
    inner = w.shape[0] // 2
    bn = _geglu_tile(2 * inner)
    half = bn // 2
    idx = []
    for t in range(2 * inner // bn):
        idx.extend(range(t * half, (t + 1) * half))
        idx.extend(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx)
    return w[idx].to(torch.float16).contiguous(), b[idx].float().contiguous()


def fold_layernorm(w, gamma, beta, bias=None):
    w32 = w.reshape(w.shape[0], -1).float()
    w16 = (w32 * gamma.float()[None, :]).to(torch.float16).contiguous()
    b2 = w32 @ beta.float()
    if bias is not None:
        b2 = b2 + bias.float()
    return w16, w16.float().sum(1).contiguous(), b2.contiguous()


def pack_geglu_ln(w, b, gamma, beta):
    w16, cs, b2 = fold_layernorm(w, gamma, beta, b)
    inner = w.shape[0] // 2
    bn = _geglu_tile(2 * inner)
    half = bn // 2
    idx = []
    for t in range(2 * inner // bn):
        idx.extend(range(t * half, (t + 1) * half))
        idx.extend(range(inner + t * half, inner + (t + 1) * half))
    idx = torch.tensor(idx)
    return w16[idx].contiguous(), b2[idx].contiguous(), cs[idx].contiguous()


def layernorm_stats(x, eps=1e-5):
    xf = x.float()
    mean = xf.mean(1)
    rstd = torch.rsqrt(xf.var(1, unbiased=False) + eps)
    return torch.stack([mean, rstd], 1).contiguous()


def _h(t):
    return t.to(torch.float16)


def linear(x, w, bias=None, res=None, geglu=False, out=None, out_f32=False, x2=None, ln=None, ln_out=False, gn_out=False, peer=None):
    # the kernel's host-side contract (gemm_tap.cu: TMA strides are 16-byte multiples)
    # (the double's own norm ops may hand back permuted views; only row-major operands carry a meaningful pitch)
    assert w.shape[1] % 8 == 0, w.shape
    for t in (x, w):
        assert t.stride(1) != 1 or t.shape[0] == 1 or t.stride(0) % 8 == 0, (t.shape, t.stride())
    a = x.float() if x2 is None else torch.cat([x.float(), x2.float()], 1)
    y = a @ w.float().t()
    if ln is not None:
        stats, colsum = ln
        y = stats[:, 1:2] * (y - stats[:, 0:1] * colsum[None, :])
    if bias is not None:
        y = y + bias
    if geglu:
        N = w.shape[0]
        bn = _geglu_tile(N)
        y = y.reshape(y.shape[0], N // bn, 2, bn // 2)
        y = (y[:, :, 0] * F.gelu(y[:, :, 1])).reshape(y.shape[0], N // 2)
    if res is not None:
        y = y + res.float()
    y = y if out_f32 else _h(y)
    if out is not None:
        out[:, :y.shape[1]].copy_(y)
        y = out
    return (y, layernorm_stats(y)) if ln_out else y


def _rows_to_nchw(x, frames, H, W):
    return x.float().reshape(frames, H, W, -1).permute(0, 3, 1, 2)


def _nchw_to_rows(y):
    return y.permute(0, 2, 3, 1).reshape(-1, y.shape[1])


def upconv3x3(x, frames, H, W, packs, bias=None):
    """the four parity sub-convolutions exactly as the kernel runs them (2x2 pre-summed taps on the small image)"""
    xi = _rows_to_nchw(x, frames, H, W)
    N = packs[0].shape[0] // 4
    out = torch.zeros((frames, N, 2 * H, 2 * W))
    for a in (0, 1):
        for b in (0, 1):
            w4 = packs[a * 2 + b].float().reshape(2, 2, N, -1).permute(2, 3, 0, 1)        # [N, K, r, c]
            xp = F.pad(xi, (1 - b, b, 1 - a, a))                                           # taps at i + r + a - 1, j + c + b - 1
            out[:, :, a::2, b::2] = F.conv2d(xp, w4)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return _h(_nchw_to_rows(out))


def conv3x3(x, frames, H, W, w9, bias=None, res=None, x2=None, bias_z_div=0, out_f32=False, out=None, gn_out=False, peer=None):
    a = x if x2 is None else torch.cat([x, x2], 1)
    N, K = w9.shape[0] // 9, w9.shape[1]
    w = w9.float().reshape(3, 3, N, K).permute(2, 3, 0, 1)
    y = _nchw_to_rows(F.conv2d(_rows_to_nchw(a, frames, H, W), w, None, padding=1))
    if bias is not None:
        y = y + (bias.repeat_interleave(bias_z_div * H * W, 0) if bias.dim() == 2 else bias)
    if res is not None:
        y = y + res.float()
    y = y if out_f32 else _h(y)
    if out is not None:
        out.copy_(y)
        return out
    return y


def conv_temporal(x, B, T, HW, w3, bias=None, res=None, gn_out=False, peer=None):
    N, K = w3.shape[0] // 3, w3.shape[1]
    w = w3.float().reshape(3, N, K).permute(1, 2, 0).reshape(N, K, 3, 1, 1)
    x5 = x.float().reshape(B, T, HW, K).permute(0, 3, 1, 2).unsqueeze(-1)
    y = F.conv3d(x5, w, bias, padding=(1, 0, 0)).squeeze(-1).permute(0, 2, 3, 1).reshape(B * T * HW, N)
    if res is not None:
        y = y + res.float()
    return _h(y)


def _attn(q, k, v, heads, scale):
    f = lambda t: t.float().reshape(t.shape[0], t.shape[1], heads, 64).permute(0, 2, 1, 3)
    s = torch.einsum("bhid,bhjd->bhij", f(q), f(k)) * scale
    o = torch.einsum("bhij,bhjd->bhid", s.softmax(-1), f(v))
    return o.permute(0, 2, 1, 3).reshape(q.shape[0], q.shape[1], heads * 64)


def flash_attn(q, k, v, B, Nq, Nk, heads, kv_shared=False, scale=0.125, out=None, accumulate=False):
    C = heads * 64
    qq = q.reshape(B, Nq, -1)[:, :, :C]
    if kv_shared:
        kk, vv = k.reshape(1, Nk, -1)[:, :, :C].expand(B, Nk, C), v.reshape(1, Nk, -1)[:, :, :C].expand(B, Nk, C)
    else:
        kk, vv = k.reshape(B, Nk, -1)[:, :, :C], v.reshape(B, Nk, -1)[:, :, :C]
    y = _attn(qq, kk, vv, heads, scale).reshape(B * Nq, C)
    if out is not None:
        out.copy_(_h(y + out.float()) if accumulate else _h(y))
        return out
    return _h(y)


def temporal_attn(q, k, v, T, sites, heads, scale=0.125, out=None):
    C = heads * 64
    tok = lambda t: t.reshape(T, sites, -1)[:, :, :C].permute(1, 0, 2)
    r = _h(_attn(tok(q), tok(k), tok(v), heads, scale).permute(1, 0, 2).reshape(T * sites, C))
    if out is not None:
        out.copy_(r)
        return out
    return r


def groupnorm(x, samples, gamma, beta, eps, silu, x2=None):
    a = x if x2 is None else torch.cat([x, x2], 1)
    rows, C = a.shape
    y = F.group_norm(a.float().reshape(samples, rows // samples, C).permute(0, 2, 1), 32, gamma, beta, eps)
    if silu:
        y = F.silu(y)
    return _h(y.permute(0, 2, 1).reshape(rows, C))


def groupnorm_stats(x, samples):
    rows, C = x.shape
    a = x.float().reshape(samples, rows // samples, 32, C // 32)
    return torch.stack([a.sum(dim=(1, 3)), (a * a).sum(dim=(1, 3))], dim=-1).contiguous()


def groupnorm_apply(x, samples, stats, stat_rows, gamma, beta, eps, silu):
    rows, C = x.shape
    n = stat_rows * (C // 32)
    mean = stats[..., 0] / n
    var = (stats[..., 1] / n - mean * mean).clamp_min(0)
    a = x.float().reshape(samples, rows // samples, 32, C // 32)
    y = (a - mean[:, None, :, None]) * torch.rsqrt(var + eps)[:, None, :, None]
    y = y.reshape(samples, rows // samples, C) * gamma + beta
    if silu:
        y = F.silu(y)
    return _h(y.reshape(rows, C))


def layernorm(x, gamma, beta, eps=1e-5):
    return _h(F.layer_norm(x.float(), (x.shape[1],), gamma, beta, eps))


def softmax_rows(x, scale):
    return _h(torch.softmax(x * scale, -1))


def upsample2x(x, N, H, W):
    return _h(_nchw_to_rows(F.interpolate(_rows_to_nchw(x, N, H, W), scale_factor=2, mode="nearest")))


def im2col_s2(x, N, H, W, pad_lo=1, pad_hi=None):
    Cc = x.shape[1]
    pad_hi = pad_lo if pad_hi is None else pad_hi
    Ho, Wo = (H + pad_lo + pad_hi - 3) // 2 + 1, (W + pad_lo + pad_hi - 3) // 2 + 1
    img = F.pad(_rows_to_nchw(x, N, H, W), (pad_lo, pad_hi, pad_lo, pad_hi))
    cols = F.unfold(img, 3, padding=0, stride=2)                                       # [N, C*9, L]: channel-major, tap-minor
    cols = cols.reshape(N, Cc, 9, Ho * Wo).permute(0, 3, 2, 1).reshape(N * Ho * Wo, 9 * Cc)
    return _h(cols), Ho, Wo


def ncthw_to_rows(x, out, c_off=0):
    Cc = x.shape[1]
    out[:, c_off:c_off + Cc] = _h(x.permute(0, 2, 3, 4, 1).reshape(-1, Cc))


def rows_to_ncthw(x, B, Cc, T, H, W):
    return x[:, :Cc].reshape(B, T, H, W, Cc).permute(0, 4, 1, 2, 3).contiguous()


def rows_f16_to_nchw(x, N, Cc, H, W):
    return x[:, :Cc].float().reshape(N, H, W, Cc).permute(0, 3, 1, 2).contiguous()


def cast_f16(x):
    return _h(x)


def gelu_f16(x):
    return _h(F.gelu(x.float()))


def add_f16(a, b):
    return _h(a.float() + b.float())


def timestep_embedding(t, dim):
    half = dim // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], -1)


def small_linear(x, w, b, silu_in=False, add=None):
    y = F.linear(F.silu(x) if silu_in else x, w, b)
    return y if add is None else y + add


def ddim_update(x, v_cond, v_uncond, noise, sc, v_uncond_img=None, cfg_img=0.0):
    from oracle import lvdm_oracle as O
    arr = np.asarray([0.0, sc["a_prev"], sc["sigma_t"], 0.0, sc["scale_t"], sc["prev_scale_t"]], dtype=np.float32)
    return O.ddim_update(x, v_cond, v_uncond, arr, sc["sqrt_ac_t"], sc["sqrt_1mac_t"], noise, sc["cfg_scale"], sc["guidance_rescale"],
                         v_uncond_img=v_uncond_img, cfg_img=cfg_img)


def install(monkeypatch):
    """Swap every public op of viewcrafter_b200.ops for the CPU double (pytest monkeypatch scope)."""
    me = sys.modules[__name__]
    for name in dir(me):
        if not name.startswith("_") and name != "install" and callable(getattr(me, name)) and hasattr(real, name):
            monkeypatch.setattr(real, name, getattr(me, name))
