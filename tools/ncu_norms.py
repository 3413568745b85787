"""ncu target: GroupNorm+SiLU (fused) at the three U-Net levels and the LayerNorm statistics pass.  ncu -k regex:gn_fused|layernorm"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
T = 25
for H, W, C in ((72, 128, 320), (36, 64, 640), (18, 32, 1280)):
    x = (torch.randn(T * H * W, C, device="cuda") * 0.7).half()
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    for _ in range(2):
        ops.groupnorm(x, T, g, b, 1e-5, True)
        ops.layernorm_stats(x)
torch.cuda.synchronize()
print("done")
