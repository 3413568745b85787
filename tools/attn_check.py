"""Correctness of the attention kernel selected by the environment (VC_ATTN_BN64) against an fp32 torch reference,
over tile counts that exercise every code path: 1 tile (one softmax group idle), odd / even counts, a masked tail, shared K/V,
the accumulate epilogue.  Prints one line per case and ATTN_CHECK_OK.  Used by tests/test_ops_gpu.py::test_attention_kernel_variants."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from viewcrafter_b200 import ops

torch.manual_seed(0)
ok = True
for (B, heads, Nq, Nk, shared, acc) in ((2, 2, 256, 128, False, False), (1, 3, 300, 300, False, False), (2, 2, 130, 77, True, False),
                                         (1, 2, 384, 256, True, True), (2, 5, 640, 1000, False, False), (1, 5, 2304, 2304, False, False),
                                         (1, 2, 1024, 9216, True, False)):
    C = heads * 64
    q = (torch.randn(B * Nq, C, device="cuda") * 0.8).half()
    kv = (torch.randn((1 if shared else B) * Nk, 2 * C, device="cuda") * 0.8).half()
    k, v = kv[:, :C], kv[:, C:]
    base = (torch.randn(B * Nq, C, device="cuda") * 0.5).half() if acc else None
    out = base.clone() if acc else None
    out = ops.flash_attn(q, k, v, B, Nq, Nk, heads, kv_shared=shared, out=out, accumulate=acc)
    worst = 0.0
    for b in range(B):
        for h in range(heads):
            qq = q[b * Nq:(b + 1) * Nq, h * 64:(h + 1) * 64].float()
            kb = 0 if shared else b
            kk, vv = k[kb * Nk:(kb + 1) * Nk, h * 64:(h + 1) * 64].float(), v[kb * Nk:(kb + 1) * Nk, h * 64:(h + 1) * 64].float()
            ref = torch.softmax(qq @ kk.t() * 0.125, -1) @ vv
            if acc:
                ref = ref + base[b * Nq:(b + 1) * Nq, h * 64:(h + 1) * 64].float()
            worst = max(worst, float((out[b * Nq:(b + 1) * Nq, h * 64:(h + 1) * 64].float() - ref).abs().max()))
    good = worst < 4e-3 and bool(torch.isfinite(out.float()).all())
    ok = ok and good
    print(f"B={B} heads={heads} Nq={Nq} Nk={Nk} shared={shared} accumulate={acc}: max err {worst:.2e} {'ok' if good else 'FAIL'}")
print("ATTN_CHECK_OK" if ok else "ATTN_CHECK_FAILED")
sys.exit(0 if ok else 1)
