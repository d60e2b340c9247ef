"""CPU check of the arithmetic and the fragment tables behind K1's integer tensor-core kernel (csrc/preprocess_mma.cuh,
area_hpass_mma_kernel; tables: csrc/preprocess.cu::get_mma_tables): no GPU, numpy only.

The kernel evaluates cv2 INTER_AREA and Pillow's horizontal pass as three chained integer matrix products whose operands
sit in mma.sync m16n8k32 / m16n8k16 register fragments.  This file restates the host-side table construction (which
weight goes to which (lane, register, byte)) and re-enacts the device data flow with an emulated MMA that honours the
PTX fragment layouts:
  stage 1  H^T = Wx^T . rows^T        A = weights [15 area columns x 64 row bytes], B = 8 source rows read as 4 x 4 bytes
  stage 2  N2^T = H^T . (2 Wy)^T + D  A = low / high byte planes of the stage-1 accumulators, repacked with PRMT pairs in
                                      the K order (2 t, 2 t + 1, 8 + 2 t, 9 + 2 t | 16 + ...) the fragments impose,
                                      K = 16 per step, planes folded after each step
  stage 3  Pillow^T = Wp^T . plane^T  three byte planes (u8, u8, s8) of the 22-bit coefficients on planar area rows,
                                      K enumerated in the order of the kernel's 64-bit loads
and compares the bytes with the oracle's restatement of cv2 + Pillow (oracle/preprocess_ref.py, itself pinned byte for
byte to the installed libraries)."""
import numpy as np
import pytest

from oracle import preprocess_ref as P

LANES = [(lane >> 2, lane & 3) for lane in range(32)]


def mma(a_frag, b_frag, c_frag, k, a_signed=False):
    """mma.sync m16n8k{k} on register fragments: a_frag [32][k/8] uint32, b_frag [32][k/16] uint32, c_frag [32][4]."""
    A = np.zeros((16, k), np.int64)
    B = np.zeros((k, 8), np.int64)
    for lane, (g, t) in enumerate(LANES):
        for r in range(k // 8):
            for j in range(4):
                v = (int(a_frag[lane][r]) >> (8 * j)) & 0xFF
                if a_signed and v >= 128:
                    v -= 256
                A[g + 8 * (r & 1), 4 * t + j + 16 * (r >> 1)] = v
        for r in range(k // 16):
            for j in range(4):
                B[4 * t + j + 16 * r, g] = (int(b_frag[lane][r]) >> (8 * j)) & 0xFF
    C = A @ B
    out = np.zeros((32, 4), np.int64)
    for lane, (g, t) in enumerate(LANES):
        out[lane] = [C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1]]
    return out + np.asarray(c_frag, np.int64)


def prmt(a, b, sel):
    src = [(a >> (8 * i)) & 0xFF for i in range(4)] + [(b >> (8 * i)) & 0xFF for i in range(4)]
    return sum(src[(sel >> (4 * i)) & 0xF] << (8 * i) for i in range(4))


def _int_taps(ssize, dsize):
    by = {}
    for d, s, w in P.area_table(ssize, dsize):
        by.setdefault(d, []).append((s, float(w)))
    den = next(c for c in range(1, 256) if all(abs(w * c - round(w * c)) < 1e-4 for v in by.values() for _, w in v))
    return [(by[d][0][0], [int(round(w * den)) for _, w in by[d]]) for d in range(dsize)], den


def put(word, j, byte):
    return word | ((byte & 0xFF) << (8 * j))


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720)])
def test_imma_formulation_equals_cv2_and_pillow_restatement(w, h):
    S = 224
    dw, dh = P.fit_size(w, h)
    xt, dx = _int_taps(w, dw)
    yt, dy = _int_taps(h, dh)
    D = dx * dy
    assert D % 2 == 1 and dx <= 255 and 2 * dy <= 255
    nw, nh = P.resize_output_size(dw, dh, S)
    left, _ = P.center_crop_box(nw, nh, S)
    bounds, kk = P.pil_coeffs(dw, nw)
    rx0 = int(min(bounds[o, 0] for o in range(left, left + S)))
    rx1 = int(max(bounds[o, 0] + bounds[o, 1] for o in range(left, left + S)))
    nx = rx1 - rx0
    sx0 = min(xt[x][0] for x in range(rx0, rx1))
    delta = 4                                              # window start 4 bytes behind a 16-byte boundary, as for 1080p
    gstart = sx0 * 3 - delta
    assert gstart >= 0

    rng = np.random.default_rng(w + h)
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frame[::7, ::2] = 255
    frame[3::11, ::3] = 0
    groups = [5, 17]                                       # two groups of 8 area rows (different row phases)
    area_want = P.inter_area_resize(frame, dw, dh)
    hp_want = P._resample_axis(area_want, nw, axis=1)[:, left:left + S]

    # ---- host tables (restated from get_mma_tables)
    ntx = (nx + 4) // 5
    a1, kb1 = [], []
    for ti in range(ntx):
        px0, px1 = rx0 + 5 * ti, min(rx0 + 5 * ti + 5, rx1)
        lastb = max(3 * (xt[x][0] + len(xt[x][1])) - gstart for x in range(px0, px1))
        kb = (3 * xt[px0][0] - gstart) & ~3
        assert kb >= 0 and lastb - kb <= 64
        frag = np.zeros((2, 32, 4), np.uint32)
        for s in range(2):
            for lane, (g, t) in enumerate(LANES):
                for r in range(4):
                    word = 0
                    for j in range(4):
                        m = g + 8 * (r & 1)
                        sb = kb + 32 * s + 16 * (r >> 1) + 4 * t + j + gstart
                        sx, sc, x = sb // 3, sb % 3, px0 + m // 3
                        if m < 15 and x < px1 and sc == m % 3 and xt[x][0] <= sx < xt[x][0] + len(xt[x][1]):
                            word = put(word, j, xt[x][1][sx - xt[x][0]])
                    frag[s, lane, r] = word
        a1.append(frag)
        kb1.append(kb)
    npt = (S + 15) // 16
    ap, kbp = [], []
    for pt in range(npt):
        o0, o1 = left + 16 * pt, min(left + 16 * pt + 16, left + S)
        lasta = max(int(bounds[o, 0] + bounds[o, 1]) - rx0 for o in range(o0, o1))
        kb = (int(bounds[o0, 0]) - rx0) & ~7
        assert kb >= 0 and lasta - kb <= 64
        frag = np.zeros((3, 2, 32, 4), np.uint32)
        for pl in range(3):
            for s in range(2):
                for lane, (g, t) in enumerate(LANES):
                    for r in range(4):
                        word = 0
                        for j in range(4):
                            o = o0 + g + 8 * (r & 1)
                            acol = kb + 32 * s + 8 * t + 4 * (r >> 1) + j + rx0
                            if o < o1 and bounds[o, 0] <= acol < bounds[o, 0] + bounds[o, 1]:
                                wv = int(kk[o, acol - bounds[o, 0]])
                                assert -(1 << 23) <= wv < (1 << 23)
                                word = put(word, j, wv >> (8 * pl))
                        frag[pl, s, lane, r] = word
        ap.append(frag)
        kbp.append(kb)

    for gi in groups:
        y0 = 8 * gi
        r_lo = yt[y0][0]
        r_hi = max(yt[y][0] + len(yt[y][1]) for y in range(y0, y0 + 8))
        nrows = r_hi - r_lo
        assert nrows <= 32
        nb = (nrows + 7) // 8
        b2 = np.zeros((32, 2), np.uint32)
        for lane, (n, t) in enumerate(LANES):
            for hh in range(2):
                word = 0
                for j in range(4):
                    sr = r_lo + 16 * hh + (2 * t + j if j < 2 else 8 + 2 * t + (j - 2))
                    y = y0 + n
                    if yt[y][0] <= sr < yt[y][0] + len(yt[y][1]):
                        word = put(word, j, 2 * yt[y][1][sr - yt[y][0]])
                b2[lane, hh] = word
        # the ring: blocks of 8 source rows; bytes past the window and rows past the group hold garbage (zero weights)
        seg = (delta + (max(xt[x][0] + len(xt[x][1]) for x in range(rx0, rx1)) - sx0) * 3 + 15) // 16 * 16
        ring = rng.integers(0, 256, (nb, 8, seg + 80), dtype=np.uint8)
        for b in range(nb):
            for r in range(8):
                if 8 * b + r < nrows:
                    ring[b, r, :seg] = frame[r_lo + 8 * b + r].reshape(-1)[gstart:gstart + seg]
        planes = rng.integers(0, 256, (3, 8, nx + 80), dtype=np.uint8)       # parked area rows, planar, garbage slack
        for ti in range(ntx):
            acc = np.full((32, 4), D, np.int64)
            pA = pB = None
            for b in range(nb):
                bfr = np.zeros((2, 32, 2), np.uint32)
                for lane, (g, t) in enumerate(LANES):
                    for q in range(4):
                        off = kb1[ti] + 4 * t + 16 * q
                        bfr[q >> 1, lane, q & 1] = int.from_bytes(ring[b, g, off:off + 4].tobytes(), "little")
                c = mma(a1[ti][0], bfr[0], np.zeros((32, 4)), 32)
                c = mma(a1[ti][1], bfr[1], c, 32)
                assert c.max() <= 255 * dx < 65536
                wA = [prmt(int(c[l][0]), int(c[l][1]), 0x5410) for l in range(32)]
                wB = [prmt(int(c[l][2]), int(c[l][3]), 0x5410) for l in range(32)]
                if b % 2 == 0 and b != nb - 1:
                    pA, pB = wA, wB
                    continue
                eA, oA = (pA, wA) if b % 2 else (wA, [0] * 32)
                eB, oB = (pB, wB) if b % 2 else (wB, [0] * 32)
                bw = [[int(b2[l][b >> 1])] for l in range(32)]
                lo = [[prmt(eA[l], oA[l], 0x6420), prmt(eB[l], oB[l], 0x6420)] for l in range(32)]
                hi = [[prmt(eA[l], oA[l], 0x7531), prmt(eB[l], oB[l], 0x7531)] for l in range(32)]
                acc = mma(lo, bw, acc, 16)
                acc = acc + (mma(hi, bw, np.zeros((32, 4)), 16) << 8)
            for lane, (g, t) in enumerate(LANES):
                for e in range(4):
                    m = g + 8 * (e >> 1)
                    x = ti * 5 + m // 3
                    if m < 15 and x < nx:
                        planes[m % 3, 2 * t + (e & 1), x] = acc[lane][e] // (2 * D)     # == umulhi(n2, magic) >> s
        got_area = np.stack([planes[c, :, :nx] for c in range(3)], axis=-1)
        assert np.array_equal(got_area, area_want[y0:y0 + 8, rx0:rx1])
        got = np.zeros((8, S, 3), np.uint8)
        for pt in range(npt):
            for ch in range(3):
                bfr = np.zeros((2, 32, 2), np.uint32)
                for lane, (g, t) in enumerate(LANES):
                    for s in range(2):
                        off = kbp[pt] + 32 * s + 8 * t
                        bfr[s, lane, 0] = int.from_bytes(planes[ch, g, off:off + 4].tobytes(), "little")
                        bfr[s, lane, 1] = int.from_bytes(planes[ch, g, off + 4:off + 8].tobytes(), "little")
                c0 = np.full((32, 4), 1 << 21, np.int64)
                c1 = np.zeros((32, 4), np.int64)
                c2 = np.zeros((32, 4), np.int64)
                for s in range(2):
                    c0 = mma(ap[pt][0][s], bfr[s], c0, 32)
                    c1 = mma(ap[pt][1][s], bfr[s], c1, 32)
                    c2 = mma(ap[pt][2][s], bfr[s], c2, 32, a_signed=True)
                acc = c0 + (c1 << 8) + (c2 << 16)
                assert np.abs(acc).max() < (1 << 31)
                for lane, (g, t) in enumerate(LANES):
                    for e in range(4):
                        ox = pt * 16 + g + 8 * (e >> 1)
                        if ox < S:
                            got[2 * t + (e & 1), ox, ch] = min(max(int(acc[lane][e]) >> 22, 0), 255)
        assert np.array_equal(got, hp_want[y0:y0 + 8])
