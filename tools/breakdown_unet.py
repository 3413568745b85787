"""Per-op time breakdown of one U-Net forward (each op timed with CUDA events + sync; development tool)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
from viewcrafter_b200.configs import UNET_PARAMS
from viewcrafter_b200.unet import UNetModel

T, H, W = 25, 72, 128
with torch.device("cuda"):
    m = UNetModel(**UNET_PARAMS)
for p in m.parameters():
    if float(p.detach().abs().max()) == 0.0:
        torch.nn.init.normal_(p, std=0.02)
m.eval()
x = torch.randn(1, 8, T, H, W, device="cuda"); ctx = torch.randn(1, 333, 1024, device="cuda")
t = torch.tensor([499], device="cuda"); fs = torch.tensor([10], device="cuda")
m(x, t, context=ctx, fs=fs); torch.cuda.synchronize()

acc = collections.defaultdict(lambda: [0, 0.0])
def wrap(name):
    fn = getattr(ops, name)
    def w(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(*a, **k); e1.record(); torch.cuda.synchronize()
        key = name
        if name == "linear":
            M, K = a[0].shape; N = a[1].shape[0]
            key = f"linear K={a[1].shape[1]:5d} N={N:5d} M={M}" + (" geglu" if k.get("geglu") else "") + (" +res" if k.get("res") is not None else "")
        elif name == "conv3x3":
            key = f"conv3x3 K={a[4].shape[1]:5d} N={a[4].shape[0]//9:5d} @{a[2]}x{a[3]}"
        elif name == "conv_temporal":
            key = f"conv_temporal C={a[0].shape[1]} HW={a[3]}"
        elif name == "flash_attn":
            key = f"flash_attn Nq={a[4]} Nk={a[5]} heads={a[6]}"
        elif name in ("groupnorm", "layernorm", "temporal_attn"):
            key = f"{name} rows={a[0].shape[0]} C={a[0].shape[1]}"
        acc[key][0] += 1; acc[key][1] += e0.elapsed_time(e1)
        return r
    setattr(ops, name, w)
for n in ("linear", "conv3x3", "conv_temporal", "flash_attn", "temporal_attn", "groupnorm", "layernorm", "upsample2x", "im2col_s2", "small_linear"):
    wrap(n)
m(x, t, context=ctx, fs=fs)
tot = sum(v[1] for v in acc.values())
print(f"sum of op times {tot:.1f} ms")
cls = collections.defaultdict(float)
for k, (n, ms) in acc.items():
    cls[k.split()[0]] += ms
print("by class:", {k: round(v, 1) for k, v in sorted(cls.items(), key=lambda kv: -kv[1])})
for k, (n, ms) in sorted(acc.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{ms:7.2f} ms n={n:3d} avg {ms/n*1e3:8.1f} us  {k}")
