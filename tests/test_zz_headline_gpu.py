"""Size-independent properties at BASELINE.json's FULL size (latent 1x4x25x72x128, full-width U-Net), where the CPU oracle
cannot run (its naive attention would need 42 GB per layer):

  * attention over N = 9216 keys: rows of softmax sum to one (V = ones gives ones) through 72 online-softmax tiles, and one
    (frame, head) matches a torch fp32 reference;
  * GroupNorm at 25x72x128x320 leaves every (sample, group) with mean 0 / variance 1;
  * one DDIM step computed as a single B=2 forward with the shared CFG prefix and the cached cross-attention K/V equals the
    step computed as two independent B=1 forwards (the reference's order, ddim.py:223-224) -- the batching / prefix / cache
    machinery changes the schedule of the work, never its result.

The file sorts last on purpose: these are the most expensive GPU tests.
"""
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600)]

T, H, W = 25, 72, 128


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viewcrafter_b200 import ops as _ops
    return _ops


def test_attention_rows_sum_to_one_and_match_fp32_at_9216_keys(ops):
    heads, N, B = 5, H * W, 2
    C = heads * 64
    g = torch.Generator().manual_seed(91)
    qkv = (torch.randn(B * N, 3 * C, generator=g) * 0.7).half().cuda()
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    qk1 = qkv.clone()
    qk1[:, 2 * C:] = 1.0                                                # V = ones, in the same [rows, 3C] layout (k and v share a row pitch)
    out1 = ops.flash_attn(qk1[:, :C], qk1[:, C:2 * C], qk1[:, 2 * C:], B, N, N, heads).float()
    assert float((out1 - 1.0).abs().max()) < 2e-3                      # fp16 P, fp32 row sum: |sum p / l - 1| is a few fp16 ulps
    out = ops.flash_attn(q, k, v, B, N, N, heads)
    for b, h in ((0, 0), (1, 4)):
        rows = slice(b * N, (b + 1) * N)
        cols = slice(h * 64, (h + 1) * 64)
        qq, kk, vv = q[rows, cols].float(), k[rows, cols].float(), v[rows, cols].float()
        ref = torch.softmax(qq @ kk.t() * 0.125, -1) @ vv
        err = (out[rows, cols].float() - ref).abs()
        assert float(err.max()) < 2e-3 and float(err.mean()) < 2e-4, (b, h, float(err.max()), float(err.mean()))


def test_groupnorm_moments_at_headline_size(ops):
    C = 320
    g = torch.Generator().manual_seed(92)
    x = (torch.randn(T * H * W, C, generator=g) * 1.7 + 0.4).half().cuda()
    y = ops.groupnorm(x, T, torch.ones(C, device="cuda"), torch.zeros(C, device="cuda"), 1e-5, False).float()
    yg = y.view(T, H * W, 32, C // 32)
    mean = yg.mean((1, 3))
    var = yg.var((1, 3), unbiased=False)
    assert float(mean.abs().max()) < 2e-3 and float((var - 1.0).abs().max()) < 5e-3, (float(mean.abs().max()), float((var - 1).abs().max()))


def test_split_groupnorm_path_equals_fused_single_launch(ops):
    """The statistics / (all-reduce) / apply form the frame-sharded 5-D GroupNorm uses across GPUs, run on one GPU at the level-1
    size, against the fused single-launch kernel (same device functions; the statistics are finalised in a different order)."""
    B, rows, C = 2, T * 36 * 64, 640
    g = torch.Generator().manual_seed(95)
    x = (torch.randn(B * rows, C, generator=g) * 1.3 - 0.2).half().cuda()
    gam, bet = (torch.rand(C, generator=g) + 0.5).cuda(), (torch.randn(C, generator=g) * 0.1).cuda()
    st = ops.groupnorm_stats(x, B)
    assert st.shape[0] == B and st.numel() == B * 64
    y_split = ops.groupnorm_apply(x, B, st, rows, gam, bet, 1e-5, True).float()
    y_fused = ops.groupnorm(x, B, gam, bet, 1e-5, True).float()
    err = (y_split - y_fused).abs()
    assert float(err.max()) <= 4e-3 + 2e-3 * float(y_fused.abs().max()) and float(err.mean()) < 2e-4, (float(err.max()), float(err.mean()))


def test_batched_cfg_step_equals_two_forwards_at_headline_size():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from viewcrafter_b200.configs import UNET_PARAMS
    from viewcrafter_b200.ddim import DDIMSampler
    from viewcrafter_b200.diffusion import LatentDiffusion
    dev = torch.device("cuda")
    torch.manual_seed(0)
    with torch.device(dev):
        model = LatentDiffusion(UNET_PARAMS, None, base_scale=0.3)
    gd = torch.Generator(device=dev).manual_seed(1)
    with torch.no_grad():
        for p in model.parameters():                                    # zero-initialised layers would make the output exactly 0
            if float(p.detach().abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=gd, device=dev) * 0.02)
    model = model.eval()
    g = torch.Generator().manual_seed(93)
    shape = (1, 4, T, H, W)
    x, cc = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
    c = {"c_crossattn": [torch.randn(1, 333, 1024, generator=g).cuda()], "c_concat": [cc]}
    uc = {"c_crossattn": [torch.randn(1, 333, 1024, generator=g).cuda()], "c_concat": [cc]}
    fs = torch.tensor([10], device=dev)
    outs = {}
    for batch_cfg in (False, True):
        smp = DDIMSampler(model, batch_cfg=batch_cfg)
        smp.make_schedule(50, "uniform_trailing", 1.0, verbose=False)
        index = 30
        step = int(smp.ddim_timesteps[index])
        ts = torch.full((1,), step, device=dev, dtype=torch.long)
        v_c, v_u = smp._apply_both(x, ts, c, uc, {"fs": fs})
        torch.manual_seed(94)
        x_prev, pred_x0 = smp.p_sample_ddim(x, c, ts, index=index, unconditional_guidance_scale=7.5, unconditional_conditioning=uc,
                                            fs=fs, guidance_rescale=0.7, _step=step)
        if batch_cfg:                                                   # a second call hits the K/V cache and must not change anything
            v_c2, _ = smp._apply_both(x, ts, c, uc, {"fs": fs})
            assert float((v_c2.float() - v_c.float()).abs().max()) <= 0.02 * float(v_c.float().std())
        outs[batch_cfg] = (v_c.float(), v_u.float(), x_prev, pred_x0)
    assert all(torch.isfinite(t).all() for pair in outs.values() for t in pair)
    assert outs[True][2].shape == shape
    std = float(outs[False][0].std())
    assert std > 1e-3                                                   # the random-weight network is not degenerate
    for i, name in enumerate(("v_cond", "v_uncond")):
        err = (outs[True][i] - outs[False][i]).abs()
        # same arithmetic per element; only the GroupNorm partial-sum split depends on the batch size, so the two runs differ by
        # fp16 rounding flips propagated through the network (the budget is the fp16-vs-fp32 budget of test_unet_gpu.py, per std)
        assert float(err.max()) <= 0.1 * std and float(err.mean()) <= 0.01 * std, (name, float(err.max()), float(err.mean()), std)
    sx = float(outs[False][2].std())
    err = (outs[True][2] - outs[False][2]).abs()                        # CFG 7.5 amplifies the U-Net differences ~16x
    assert float(err.mean()) <= 0.05 * sx, (float(err.mean()), sx)
