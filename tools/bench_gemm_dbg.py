import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
def rnd(*s): return (torch.randn(*s, device="cuda") * 0.05).half()
T = 25
out = []
for name, H, W, Ci, Co in (("conv l0 320->320 (BN160)", 72, 128, 320, 320), ("conv l2 1280->1280 (BN256)", 18, 32, 1280, 1280)):
    x, w9 = rnd(T * H * W, Ci), rnd(9 * Co, Ci)
    out.append((name, t(lambda: ops.conv3x3(x, T, H, W, w9))))
x, w = rnd(T * 9216, 320), rnd(960, 320)
out.append(("qkv l0 320->960", t(lambda: ops.linear(x, w))))
print("VC_GEMM_DEBUG=%s PAIR=%s : " % (os.environ.get("VC_GEMM_DEBUG", "0"), os.environ.get("VC_GEMM_PAIR", "1")) + " | ".join(f"{n} {us:7.1f} us" for n, us in out))
