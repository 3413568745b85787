"""Frame-sharded multi-GPU execution of the U-Net forward (one process per GPU, torch.distributed / NCCL over NVLink).

SURVEY.md 8(e): every spatial op (2-D convs, per-frame GroupNorm, SpatialTransformer incl. attention, Down/Upsample)
is independent per frame, so rank r owns a contiguous range of the T frames.  The ops that couple frames --
TemporalTransformer (17 per forward) and TemporalConvBlock (22 per forward), including their 5-D GroupNorm statistics
-- run in the transposed "site" layout: every rank holds ALL T frames of H*W/P pixels (H*W is divisible by 8 at every
level), which is perfectly balanced.  The two layouts are exchanged with ONE uneven all-to-all each way
(NCCL all_to_all_single over NVSwitch; volume per rank = activation_bytes * (P-1)/P^2) and the 5-D GroupNorm adds a
[B,32,2] all-reduce.  The reference has no multi-GPU path for this; this is new functionality.
"""
from __future__ import annotations

from typing import List

import torch


def frame_ranges(T: int, world: int) -> List[tuple]:
    """Contiguous, as-even-as-possible split of T frames: 25 over 8 -> 4,3,3,3,3,3,3,3."""
    base, extra = divmod(T, world)
    out, f = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((f, f + n))
        f += n
    return out


class FrameComm:
    def __init__(self, dist, rank: int, world: int, group=None):
        self.dist, self.rank, self.world, self.group = dist, rank, world, group
        self.T = None
        self.ranges = None
        self.bytes_moved = 0            # all-to-all payload sent by this rank (for the bench report)

    def __bool__(self):
        return self.world > 1

    def bind(self, T: int):
        self.T = T
        self.ranges = frame_ranges(T, self.world)
        return self.ranges[self.rank]

    # -- layout transposes -----------------------------------------------------------------------
    def to_sites(self, h: torch.Tensor, B: int, HW: int) -> torch.Tensor:
        """[(b, t_local, hw), C] -> [(b, t_all, hw_local), C]."""
        P, C = self.world, h.shape[1]
        assert HW % P == 0, f"H*W={HW} must be divisible by the world size {P}"
        HWl = HW // P
        Tl = self.ranges[self.rank][1] - self.ranges[self.rank][0]
        send = h.view(B, Tl, P, HWl, C).permute(2, 0, 1, 3, 4).contiguous().view(P * B * Tl * HWl, C)
        out_rows = [B * (f1 - f0) * HWl for f0, f1 in self.ranges]
        recv = torch.empty((sum(out_rows), C), device=h.device, dtype=h.dtype)
        self.dist.all_to_all_single(recv, send, output_split_sizes=out_rows, input_split_sizes=[B * Tl * HWl] * P, group=self.group)
        self.bytes_moved += send.numel() * send.element_size() * (P - 1) // P
        if B == 1:
            return recv                                                   # chunks arrive in frame order already
        parts = [c.view(B, -1, HWl, C) for c in torch.split(recv, out_rows, 0)]
        return torch.cat(parts, dim=1).reshape(B * self.T * HWl, C)

    def to_frames(self, t: torch.Tensor, B: int, HW: int) -> torch.Tensor:
        """[(b, t_all, hw_local), C] -> [(b, t_local, hw), C]."""
        P, C = self.world, t.shape[1]
        HWl = HW // P
        Tl = self.ranges[self.rank][1] - self.ranges[self.rank][0]
        in_rows = [B * (f1 - f0) * HWl for f0, f1 in self.ranges]
        if B == 1:
            send = t
        else:
            t4 = t.view(B, self.T, HWl, C)
            send = torch.cat([t4[:, f0:f1].reshape(-1, C) for f0, f1 in self.ranges], 0)
        recv = torch.empty((P * B * Tl * HWl, C), device=t.device, dtype=t.dtype)
        self.dist.all_to_all_single(recv, send.contiguous(), output_split_sizes=[B * Tl * HWl] * P, input_split_sizes=in_rows, group=self.group)
        self.bytes_moved += (send.numel() - in_rows[self.rank] * C) * send.element_size()
        return recv.view(P, B, Tl, HWl, C).permute(1, 2, 0, 3, 4).contiguous().view(B * Tl * HW, C)

    def all_reduce(self, t: torch.Tensor):
        self.dist.all_reduce(t, group=self.group)

    def gather_frames(self, y_local: torch.Tensor, T: int) -> torch.Tensor:
        """[B,C,T_local,H,W] per rank -> the full [B,C,T,H,W] on every rank (3.7 MB at the headline size)."""
        B, C, _, H, W = y_local.shape
        full = torch.zeros((B, C, T, H, W), device=y_local.device, dtype=y_local.dtype)
        f0, f1 = self.ranges[self.rank]
        full[:, :, f0:f1] = y_local
        self.dist.all_reduce(full, group=self.group)
        return full


class CfgComm:
    """Classifier-free-guidance split: the conditional and unconditional U-Net forwards of a DDIM step are independent
    (ddim.py:223-224), so the first half of the ranks computes `cond`, the second half `uncond`, and rank i swaps its
    3.7 MB prediction with rank i + world/2 through a 2-rank all-reduce."""

    def __init__(self, dist, branch: int, pair_group):
        self.dist, self.branch, self.pair_group = dist, branch, pair_group

    def exchange(self, v_mine: torch.Tensor):
        buf = torch.zeros((2, *v_mine.shape), device=v_mine.device, dtype=v_mine.dtype)
        buf[self.branch] = v_mine
        self.dist.all_reduce(buf, group=self.pair_group)
        return buf[0], buf[1]


def shard_model(model, dist, rank: int, world: int, cfg_split: bool = True):
    """Distribute the denoise step over `world` ranks (weights stay replicated: 2.9 GB fp16 per GPU).

    world even and cfg_split: 2-way CFG split x (world/2)-way frame sharding -- e.g. 8 GPUs = 2 x 4 with frames 7/6/6/6
    (ideal 7.1x) instead of 8-way frames 4/3x7 (ideal 6.25x).  Otherwise pure frame sharding.
    Every rank must call this (it creates process groups collectively).  Returns the FrameComm (or None)."""
    unet = model.model.diffusion_model if hasattr(model, "model") else model
    if cfg_split and world % 2 == 0 and hasattr(model, "model"):
        P = world // 2
        frame_groups = [dist.new_group(list(range(b * P, (b + 1) * P))) for b in range(2)]
        pair_groups = [dist.new_group([i, i + P]) for i in range(P)]
        branch, r = rank // P, rank % P
        model._cfg = CfgComm(dist, branch, pair_groups[r])
        unet._comm = FrameComm(dist, r, P, frame_groups[branch]) if P > 1 else None
        return unet._comm
    unet._comm = FrameComm(dist, rank, world, None)
    return unet._comm
