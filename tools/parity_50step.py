"""SURVEY.md 8(d) config 2/3 parity over a WHOLE sampling run: x_0 after 50 DDIM steps (and after 1 step) of the CUDA path vs the
oracle in fp32 on the same GPU, with identical x_T and identical per-step noise tensors on both sides, next to
E_ref = the oracle under torch.autocast(fp16) vs the oracle in fp32 over the same 50 steps (what the reference's own fp16 mode
drifts by).  Full-width U-Net, latent 25x4x40x64 (ViewCrafter_25_512; the fp32 oracle needs ~2 s per forward there),
cfg 7.5, guidance rescale 0.7, eta 1, uniform_trailing.

    python tools/parity_50step.py [--steps 50] [--H 40 --W 64] > gpurun_out/parity_50step.json
"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=50)
ap.add_argument("--H", type=int, default=40)
ap.add_argument("--W", type=int, default=64)
ap.add_argument("--base-scale", type=float, default=0.7)
args = ap.parse_args()

from oracle import lvdm_oracle as O
from viewcrafter_b200.configs import UNET_PARAMS
from viewcrafter_b200.ddim import DDIMSampler
from viewcrafter_b200.diffusion import LatentDiffusion

dev = torch.device("cuda")
torch.manual_seed(0)
with torch.device(dev):
    model = LatentDiffusion(UNET_PARAMS, None, base_scale=args.base_scale)
gd = torch.Generator(device=dev).manual_seed(1)
with torch.no_grad():
    for p in model.parameters():
        if float(p.detach().abs().max()) == 0.0:
            p.copy_(torch.randn(p.shape, generator=gd, device=dev) * 0.02)
model = model.eval()
unet = model.model.diffusion_model
unet.enable_cuda_graph()
sd = {k: v.detach() for k, v in unet.state_dict().items()}
T, H, W, S = 25, args.H, args.W, args.steps
g = torch.Generator().manual_seed(2)
shape = (1, 4, T, H, W)
x_T, cc = torch.randn(shape, generator=g).cuda(), torch.randn(shape, generator=g).cuda()
ctx_c, ctx_u = torch.randn(1, 333, 1024, generator=g).cuda(), torch.randn(1, 333, 1024, generator=g).cuda()
fs = torch.tensor([10], device=dev)
c = {"c_crossattn": [ctx_c], "c_concat": [cc]}
uc = {"c_crossattn": [ctx_u], "c_concat": [cc]}

t0 = time.time()
torch.manual_seed(7)
smp = DDIMSampler(model, batch_cfg=True)
ours, inter = smp.sample(S=S, batch_size=1, shape=shape[1:], conditioning=c, eta=1.0, verbose=False, x_T=x_T,
                         unconditional_guidance_scale=7.5, unconditional_conditioning=uc, fs=fs, timestep_spacing="uniform_trailing",
                         guidance_rescale=0.7, log_every_t=1)
torch.cuda.synchronize()
t_ours = time.time() - t0
torch.manual_seed(7)
noises = [torch.randn(shape, device=dev) for _ in range(S)]            # the draws p_sample_ddim made, in order

sched = {k: v.cuda() for k, v in O.model_schedule(base_scale=args.base_scale).items()}


def model_fn(x, ts, cond):
    return O.unet_forward(sd, torch.cat([x, cc], 1), ts, cond, fs)


def run(autocast):
    with torch.no_grad():
        if autocast:
            with torch.autocast("cuda", dtype=torch.float16):
                return O.ddim_sample(lambda x, ts, cond: model_fn(x, ts, cond).float(), sched, shape, S, ctx_c, ctx_u, x_T, noises,
                                     eta=1.0, cfg_scale=7.5, guidance_rescale=0.7, log_every_t=1)
        with O.exact_fp32():
            return O.ddim_sample(model_fn, sched, shape, S, ctx_c, ctx_u, x_T, noises, eta=1.0, cfg_scale=7.5, guidance_rescale=0.7,
                                 log_every_t=1)


t0 = time.time(); ref16, inter16 = run(True); torch.cuda.synchronize(); t16 = time.time() - t0
t0 = time.time(); ref32, inter32 = run(False); torch.cuda.synchronize(); t32 = time.time() - t0


def cmp(a, b):
    d = (a.float() - b.float()).abs()
    return {"max": float(d.max()), "mean": float(d.mean())}


res = {"config": "latent 1x4x%dx%dx%d, S=%d, cfg 7.5, rescale 0.7, eta 1, uniform_trailing, base_scale %.1f" % (T, H, W, S, args.base_scale),
       "x0_std": float(ref32.std()),
       "after_1_step": {"ours_vs_fp32": cmp(inter["x_inter"][1], inter32["x_inter"][1]), "e_ref_fp16_autocast_vs_fp32": cmp(inter16["x_inter"][1], inter32["x_inter"][1])},
       "after_%d_steps" % S: {"ours_vs_fp32": cmp(ours, ref32), "e_ref_fp16_autocast_vs_fp32": cmp(ref16, ref32), "ours_vs_fp16_autocast": cmp(ours, ref16)},
       "trajectory_max_err_every_10": [{"step": i, "ours": cmp(inter["x_inter"][i], inter32["x_inter"][i])["max"],
                                        "e_ref": cmp(inter16["x_inter"][i], inter32["x_inter"][i])["max"]} for i in range(0, len(inter32["x_inter"]), 10)],
       "seconds": {"ours_%d_steps" % S: t_ours, "oracle_fp16_autocast": t16, "oracle_fp32": t32}}
print(json.dumps(res, indent=1))
