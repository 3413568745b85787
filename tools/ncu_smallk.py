"""ncu target: the three short-K (K = 320) level-0 linears.  ncu -k regex:gemm_tap --launch-skip 3 -c 3 ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
M = 25 * 72 * 128
x = (torch.randn(M, 320, device="cuda") * 0.5).half()
res = (torch.randn(M, 320, device="cuda") * 0.5).half()
w1 = (torch.randn(320, 320, device="cuda") * 0.05).half()
wq = (torch.randn(960, 320, device="cuda") * 0.05).half()
wg, bg = ops.pack_geglu((torch.randn(2560, 320, device="cuda") * 0.05), torch.zeros(2560, device="cuda"))
b = torch.zeros(320, device="cuda")
for _ in range(2):
    ops.linear(x, w1, bias=b, res=res)
    ops.linear(x, wq)
    ops.linear(x, wg, bias=bg, geglu=True)
torch.cuda.synchronize()
print("done")
