"""VAE decode throughput at the headline size: 25 latent frames 72x128 -> 576x1024, per-frame
(the reference's perframe_ae loop, ddpm3d.py:646-671) vs batched decode.  Random-init full-width decoder."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200.autoencoder import AutoencoderKL
from viewcrafter_b200.configs import VAE_DDCONFIG

torch.manual_seed(0)
with torch.device("cuda"):
    vae = AutoencoderKL(VAE_DDCONFIG, None, 4).eval()
z = torch.randn(25, 4, 72, 128, device="cuda")


def timed(fn, reps=2):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


x = torch.rand(25, 3, 576, 1024, device="cuda") * 2 - 1
with torch.no_grad():
    te = timed(lambda: [vae.encode(x[i:i + 1]) for i in range(25)])
    print(f"per-frame encode (25 calls): {te * 1e3:.1f} ms = {25 / te:.1f} frames/s  ({2.6 * 25 / te:.0f} TFLOP/s at ~2.6 TFLOP/frame)")
    t1 = timed(lambda: [vae.decode(z[i:i + 1]) for i in range(25)])
    print(f"per-frame decode (25 calls): {t1 * 1e3:.1f} ms = {25 / t1:.1f} frames/s  ({5.754 * 25 / t1:.0f} TFLOP/s at 5.754 TFLOP/frame)")
    for nb in (5, 25):
        try:
            t = timed(lambda: [vae.decode(z[i:i + nb]) for i in range(0, 25, nb)])
            print(f"batched decode ({nb} frames/call): {t * 1e3:.1f} ms = {25 / t:.1f} frames/s  ({5.754 * 25 / t:.0f} TFLOP/s), peak mem {torch.cuda.max_memory_allocated() / 2**30:.1f} GB")
        except Exception as e:      # noqa
            print(f"batched decode ({nb}): {type(e).__name__}: {str(e)[:200]}")
