"""Limiter matrix for the short-K (K = 320) level-0 linears: run once per VC_GEMM_DEBUG value (bit 0 skip MMA, bit 1 skip
TMA, bit 2 skip epilogue) and compare.  Usage: for d in 0 1 2 3 4 5 6 7; do VC_GEMM_DEBUG=$d python tools/bench_gemm_smallk_dbg.py; done"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops


def t(fn, reps=10):
    fn(); fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def rnd(*s):
    return (torch.randn(*s, device="cuda") * 0.05).half()


M = 25 * 9216
x, x4 = rnd(M, 320), rnd(M, 1280)
w1, w3, w4 = rnd(320, 320), rnd(960, 320), rnd(320, 1280)
wg, bg = ops.pack_geglu(torch.randn(2560, 320, device="cuda") * 0.05, torch.zeros(2560, device="cuda"))
b1 = torch.zeros(320, device="cuda")
res = rnd(M, 320)
out = [
    ("lin+res", t(lambda: ops.linear(x, w1, bias=b1, res=res))),
    ("lin", t(lambda: ops.linear(x, w1, bias=b1))),
    ("qkv", t(lambda: ops.linear(x, w3))),
    ("geglu", t(lambda: ops.linear(x, wg, bias=bg, geglu=True))),
    ("ff2+res", t(lambda: ops.linear(x4, w4, bias=b1, res=res))),
]
print("DEBUG=%s : " % os.environ.get("VC_GEMM_DEBUG", "0") + " | ".join(f"{n} {us:7.1f}" for n, us in out))
