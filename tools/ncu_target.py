"""Short target for `ncu --set full`: a few launches of the dominant kernels at their headline shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
T, H, W, C = 25, 72, 128, 320
M = T * H * W
which = sys.argv[1] if len(sys.argv) > 1 else "all"
REPS = int(os.environ.get("VC_NCU_REPS", "3"))      # launches of each op (ncu --set full replays every launch ~40 times: 1 is enough there)
x = (torch.randn(M, C, device="cuda") * 0.5).half()
if which in ("gemm", "all"):
    w9 = (torch.randn(9 * C, C, device="cuda") * 0.02).half()
    b = torch.zeros(C, device="cuda")
    for _ in range(REPS):
        ops.conv3x3(x, T, H, W, w9)                     # exactly bench.py's `roofline` launch
    ops.conv3x3(x, T, H, W, w9, bias=b, res=x)          # the ResBlock form (bias + residual epilogue)
if which in ("attn", "all"):
    qkv = (torch.randn(M, 3 * C, device="cuda") * 0.5).half()
    for _ in range(REPS):
        ops.flash_attn(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], T, H * W, H * W, 5)
if which in ("norm", "all"):
    g, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(REPS):
        ops.groupnorm(x, T, g, be, 1e-5, True)
        ops.layernorm_stats(x)
if which in ("gnparts", "all"):
    # GroupNorm whose statistics come from the producing conv's epilogue: conv with gn_out, then finalize + one-pass normalise
    ops.GN_FROM_PRODUCER = 2
    ops.GN_PARTS_MIN_MB = 0.0
    w9 = (torch.randn(9 * C, C, device="cuda") * 0.02).half()
    g, be = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(REPS):
        y = ops.conv3x3(x, T, H, W, w9, gn_out=True)
        ops.groupnorm(y, T, g, be, 1e-5, True)
        ops.groupnorm(y, 1, g, be, 1e-5, True)
if which in ("lin", "all"):
    # level-0 transformer linears with the LayerNorm folded into the epilogue: qkv (320->960) and GEGLU (320->2560)
    g32, b32 = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    wq, csq, bq = ops.fold_layernorm(torch.randn(3 * C, C, device="cuda") * 0.05, g32, b32)
    wg, bg, csg = ops.pack_geglu_ln(torch.randn(8 * C, C, device="cuda") * 0.05, torch.zeros(8 * C, device="cuda"), g32, b32)
    st = ops.layernorm_stats(x)
    for _ in range(REPS):
        ops.linear(x, wq, bias=bq, ln=(st, csq))
        ops.linear(x, wg, bias=bg, geglu=True, ln=(st, csg))
if which in ("tattn", "all"):
    qkv = (torch.randn(M, 3 * C, device="cuda") * 0.5).half()
    for _ in range(REPS):
        ops.temporal_attn(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], T, H * W, 5)
torch.cuda.synchronize()
print("done")
