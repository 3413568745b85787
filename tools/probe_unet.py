"""Time one full-size U-Net forward (random init, synthetic inputs) -- development probe, not the bench."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200.unet import UNetModel
from viewcrafter_b200.configs import UNET_PARAMS as UNET_KW

T, H, W = [int(a) for a in (sys.argv[1:4] if len(sys.argv) > 3 else (25, 72, 128))]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
torch.manual_seed(0)
t0 = time.time()
with torch.device("cuda"):
    m = UNetModel(**UNET_KW)
for p in m.parameters():
    if float(p.abs().max()) == 0.0:
        torch.nn.init.normal_(p, std=0.02)
m.eval()
print("build %.1fs" % (time.time() - t0), flush=True)
x = torch.randn(1, 8, T, H, W, device="cuda")
ctx = torch.randn(1, 333, 1024, device="cuda")
t = torch.tensor([499], device="cuda"); fs = torch.tensor([10], device="cuda")
y = m(x, t, context=ctx, fs=fs)
torch.cuda.synchronize()
print("first forward done; out std %.4f finite %s; peak mem %.1f GB" % (float(y.std()), bool(torch.isfinite(y).all()), torch.cuda.max_memory_allocated() / 2**30), flush=True)
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0 = time.time()
    e0.record(); y = m(x, t, context=ctx, fs=fs); e1.record()
    host = time.time() - h0
    torch.cuda.synchronize()
    print("forward %dx%dx%d: %.1f ms device, %.1f ms host-issue" % (T, H, W, e0.elapsed_time(e1), host * 1e3), flush=True)
