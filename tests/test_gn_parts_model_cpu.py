"""Executable specification of the GroupNorm partial sums a GEMM epilogue leaves (csrc/gemm_common.cuh: gn_part_accumulate) and of how
norm.cu: gn_part_finalize_kernel folds them into the groups of a consumer: every 32-column chunk of a 32-row block is cut into 4 pieces
at multiples of `sub` channels; group g of a GroupNorm(32) over concat(x1, x2) is the union of whole sub-groups of its sources.  numpy
transcription of both index computations against the plain per-group sums."""
import numpy as np
import pytest


def epilogue_records(y, sub):
    """y [rows, N] -> records [rows/32, N/32, 4, 2] exactly as the epilogue cuts them (rows % 32 == 0 here)."""
    rows, N = y.shape
    hp = sub // 2
    rec = np.zeros((rows // 32, N // 32, 4, 2))
    for c in range(N // 32):
        nb = 32 * c
        o = nb % sub
        b1 = (sub - o) // 2 if sub == 10 else 4
        for e in range(16):                                    # column pairs of the chunk
            piece = 0 if e < b1 else 1 if e < b1 + hp else 2 if e < b1 + 2 * hp else 3
            cols = y[:, nb + 2 * e: nb + 2 * e + 2].reshape(rows // 32, 32, 2)
            rec[:, c, piece, 0] += cols.sum((1, 2))
            rec[:, c, piece, 1] += (cols ** 2).sum((1, 2))
    return rec


def finalize(rec, sub, C_src, cg, c_off):
    """records of ONE source -> [32, 2] group sums of the consumer (the tid < 64 part of gn_part_finalize_kernel)."""
    red = rec.sum(0).reshape(-1, 2)                            # [cols = chunks * 4, 2]
    out = np.zeros((32, 2))
    for grp in range(32):
        ch_lo, ch_hi = max(grp * cg - c_off, 0), min((grp + 1) * cg - c_off, C_src)
        if ch_hi <= ch_lo:
            continue
        for sg in range(ch_lo // sub, ch_hi // sub):
            for c in range((sg * sub) >> 5, ((sg * sub + sub - 1) >> 5) + 1):
                k = sg - (c * 32) // sub
                if 0 <= k < 4:
                    out[grp] += red[c * 4 + k]
    return out


@pytest.mark.parametrize("C1,C2,sub", [(320, 0, 10), (640, 0, 10), (1280, 0, 10), (320, 320, 10), (640, 320, 10), (640, 640, 10), (1280, 640, 10),
                                       (1280, 1280, 10), (256, 0, 8), (512, 0, 8), (512, 256, 8)])
def test_pieces_reassemble_into_every_consumers_groups(C1, C2, sub):
    rng = np.random.default_rng(C1 + C2)
    rows = 64
    x1 = rng.standard_normal((rows, C1))
    x2 = rng.standard_normal((rows, C2)) if C2 else None
    C = C1 + C2
    cg = C // 32
    assert cg % sub == 0 and C1 % sub == 0
    got = finalize(epilogue_records(x1, sub), sub, C1, cg, 0)
    if C2:
        got = got + finalize(epilogue_records(x2, sub), sub, C2, cg, C1)
    full = x1 if x2 is None else np.concatenate([x1, x2], 1)
    want = np.stack([full.reshape(rows, 32, cg).sum((0, 2)), (full ** 2).reshape(rows, 32, cg).sum((0, 2))], 1)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-9)
    # every chunk has exactly four non-empty pieces (what makes the fixed [chunk][4] record layout possible)
    rec = epilogue_records(np.ones((32, C1)), sub)
    assert (rec[..., 0] > 0).all()
