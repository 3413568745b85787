"""Executable specification of the layout switch fused into the GEMM epilogue (csrc/gemm_common.cuh: peer_scatter32, gemm_tap.cu: GemmPeer
maps): a Python transcription of the per-patch routing -- 32-row epilogue patches, one clipped 3-D TMA store per row segment into the
per-rank destination maps -- applied to index tensors and compared with the frame <-> site permutation that parallel.FrameComm defines
(to_sites / to_frames).  Covers the producer geometries (3x3 conv tiles, temporal conv, plain linear incl. several batch elements in one
row matrix), CTA-pair padding tiles, 2 and 4 ranks, uneven frame ranges, and rank / frame straddling patches (the clipping model; on
hardware those shapes stay on the separate exchange kernel -- TMA stores with negative start coordinates fault)."""
import numpy as np
import pytest

from viewcrafter_b200.parallel import frame_ranges


def conv_box(H, W):
    if W >= 128 or 128 % W != 0:
        return 128, 1
    return W, 128 // W


def patches(X, Y, Z, bx, by, pair):
    """(wx, wy, wz, [row ids of the 32 patch rows or -1]) for every epilogue warp patch of every m-tile (x fastest, then y, then z)."""
    tx, ty = -(-X // bx), -(-Y // by)
    m_tiles = tx * ty * Z
    n = (m_tiles + 1) // 2 * 2 if pair else m_tiles
    bw = min(bx, 32)
    for m in range(n):
        q1, txi = divmod(m, tx)
        z, tyi = divmod(q1, ty)
        x0, y0 = txi * bx, tyi * by
        for quad in range(4):
            R0 = quad * 32
            wx, wy = x0 + (R0 & (bx - 1)), y0 + R0 // bx
            rows = []
            for i in range(32):                       # patch row i = pixel (wx + i % bw, wy + i // bw) of slab z
                x, y = wx + i % bw, wy + i // bw
                rows.append((z * Y + y) * X + x if (x < X and y < Y and z < Z) else -1)
            yield wx, wy, z, rows


def scatter(mode, P, me, B, T, HW, ranges, X, Y, Z, bx, by, pair, src_ids, dst):
    """Route the patches of rank `me` (device code transcription).  dst[q] = the destination map of rank q: array [c2, HWl]."""
    HWl = HW // P
    f0 = [r[0] for r in ranges] + [T]
    Tl_me = ranges[me][1] - ranges[me][0]
    rps = HW if mode == 1 else T * HWl
    wrap = (Y == 1 and Z == 1)
    for wx, wy, wz, rows in patches(X, Y, Z, bx, by, pair):
        if wz >= Z:
            continue
        lin, slab = wy * X + wx, wz
        if wrap:
            slab, lin = divmod(lin, rps)
        rem, i0 = 32, 0
        while rem > 0:
            if lin >= rps:
                if not wrap:
                    break
                lin -= rps; slab += 1
            if mode == 1:
                q, s = divmod(lin, HWl)
                b, tl = divmod(slab, Tl_me)
                c2 = b * T + f0[me] + tl
            else:
                tt, s = divmod(lin, HWl)
                q = 0
                while q + 1 < P and tt >= f0[q + 1]:
                    q += 1
                c2 = slab * (f0[q + 1] - f0[q]) + tt - f0[q]
            ln = min(rem, HWl - s)
            # one 3-D TMA store of the whole 32-row tile at (dim1 = s - i0, dim2 = c2): rows outside [0, HWl) x [0, dims2) are clipped
            if 0 <= c2 < dst[q].shape[0]:
                for i in range(32):
                    d1 = s - i0 + i
                    if 0 <= d1 < HWl:
                        assert rows[i] >= 0, "a padding row was routed into a valid destination row"
                        dst[q][c2, d1] = src_ids[rows[i]]
            i0 += ln; lin += ln; rem -= ln


@pytest.mark.parametrize("P,T,B,H,W,geom", [
    (2, 25, 1, 72, 128, "conv"), (2, 25, 1, 36, 64, "conv"), (2, 25, 1, 18, 32, "conv"), (2, 5, 2, 16, 16, "conv"),
    (4, 25, 1, 36, 64, "conv"), (4, 7, 1, 16, 32, "linear"), (2, 5, 2, 16, 16, "linear"), (4, 25, 1, 72, 128, "linear"),
    (2, 25, 1, 9, 16, "conv"), (4, 25, 1, 9, 16, "linear"), (4, 25, 1, 18, 32, "conv"),            # straddling patches (clipping model)
])
def test_frames_to_sites_routing_equals_the_layout_permutation(P, T, B, H, W, geom):
    HW, HWl = H * W, H * W // P
    ranges = frame_ranges(T, P)
    dst = [np.full((B * T, HWl), -1, dtype=np.int64) for _ in range(P)]
    want = [np.full((B * T, HWl), -1, dtype=np.int64) for _ in range(P)]
    for me in range(P):
        f_lo, f_hi = ranges[me]
        Tl = f_hi - f_lo
        M = B * Tl * HW
        ids = (me << 40) + np.arange(M, dtype=np.int64)                       # unique id of every source row
        for b in range(B):
            for tl in range(Tl):
                for hw in range(HW):
                    want[hw // HWl][b * T + f_lo + tl, hw % HWl] = ids[(b * Tl + tl) * HW + hw]
        if geom == "conv":
            bx, by = conv_box(H, W)
            scatter(1, P, me, B, T, HW, ranges, W, H, B * Tl, bx, by, True, ids, dst)
        else:
            scatter(1, P, me, B, T, HW, ranges, M, 1, 1, 128, 1, True, ids, dst)
    for q in range(P):
        assert np.array_equal(dst[q], want[q]), f"rank {q}"


@pytest.mark.parametrize("P,T,B,HW,geom", [
    (2, 25, 1, 9216, "tconv"), (2, 25, 1, 2304, "tconv"), (2, 25, 1, 2304, "linear"), (2, 5, 2, 256, "tconv"), (2, 5, 2, 256, "linear"),
    (4, 25, 1, 2304, "tconv"), (4, 7, 2, 512, "linear"), (2, 25, 1, 144, "tconv"), (4, 25, 1, 576, "linear"), (4, 25, 1, 144, "tconv"),
])
def test_sites_to_frames_routing_equals_the_layout_permutation(P, T, B, HW, geom):
    HWl = HW // P
    ranges = frame_ranges(T, P)
    tls = [r[1] - r[0] for r in ranges]
    # destination of rank q: its frames buffer [(b, t_local), HW]; the map of SENDER `me` is the column window [me*HWl, (me+1)*HWl)
    full = [np.full((B * tls[q], HW), -1, dtype=np.int64) for q in range(P)]
    want = [np.full((B * tls[q], HW), -1, dtype=np.int64) for q in range(P)]
    for me in range(P):
        M = B * T * HWl
        ids = (me << 40) + np.arange(M, dtype=np.int64)
        for b in range(B):
            for t in range(T):
                q = next(r for r in range(P) if ranges[r][0] <= t < ranges[r][1])
                for s in range(HWl):
                    want[q][b * tls[q] + t - ranges[q][0], me * HWl + s] = ids[(b * T + t) * HWl + s]
        dst = [full[q][:, me * HWl:(me + 1) * HWl] for q in range(P)]
        if geom == "tconv":
            scatter(2, P, me, B, T, HW, ranges, T * HWl, 1, B, 128, 1, True, ids, dst)
        else:
            scatter(2, P, me, B, T, HW, ranges, M, 1, 1, 128, 1, True, ids, dst)
    for q in range(P):
        assert np.array_equal(full[q], want[q]), f"rank {q}"
