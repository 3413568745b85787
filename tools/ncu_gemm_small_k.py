"""ncu target: the K=320 GEMMs of the level-0 transformer blocks (epilogue-heavy) + the big 3x3 conv."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
M = 25 * 72 * 128
x = (torch.randn(M, 320, device="cuda") * 0.5).half()
wq = (torch.randn(960, 320, device="cuda") * 0.05).half()
wg, bg = ops.pack_geglu((torch.randn(2560, 320, device="cuda") * 0.05), torch.zeros(2560, device="cuda"))
w9 = (torch.randn(9 * 320, 320, device="cuda") * 0.02).half()
b = torch.zeros(320, device="cuda")
for _ in range(2):
    ops.linear(x, wq)
    ops.linear(x, wg, bias=bg, geglu=True)
    ops.conv3x3(x, 25, 72, 128, w9, bias=b, res=x)
torch.cuda.synchronize()
print("done")
