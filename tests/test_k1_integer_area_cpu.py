"""CPU check of the arithmetic behind K1's integer-exact, vertical-first area kernel (csrc/preprocess.cu,
area_hpass_vfirst_kernel): no GPU, numpy only.

(1) For the 1080p and 720p geometries every cv2 INTER_AREA weight is a multiple of 1/15 (1/5), so an output sample is
    N / (Dx*Dy) with N an integer, and rint() of OpenCV's fp32 evaluation equals floor((2N + D) / 2D) (odd D: no ties).
(2) The kernel accumulates DOWN the rows first, with the even and the odd bytes of every 32-bit word in 16-bit lanes,
    parks the lanes as an even-byte stream L and an odd-byte stream H, and combines the 15 sums under an area column
    with IDP.2A from two realigned streams X (bytes s0, s0+2, ...) and Y (bytes s0+1, s0+3, ...).  This file re-enacts
    that index arithmetic (start byte, parity, funnel shift, the 15 (register, lane, tap) triples and their packed
    byte weights) and compares the result with the oracle's restatement of cv2 (oracle/preprocess_ref.py), which is
    itself pinned byte for byte against the installed OpenCV."""
import numpy as np
import pytest

from oracle import preprocess_ref as P

# (stream, element, tap) per channel, exactly the order of the kernel's dp2a chain: byte offset o = 3*tap + channel is
# X[o/2] when o is even and Y[(o-1)/2] when it is odd
CHAIN = {
    0: [("x", 0, 0), ("x", 3, 2), ("x", 6, 4), ("y", 1, 1), ("y", 4, 3)],
    1: [("x", 2, 1), ("x", 5, 3), ("y", 0, 0), ("y", 3, 2), ("y", 6, 4)],
    2: [("x", 1, 0), ("x", 4, 2), ("x", 7, 4), ("y", 2, 1), ("y", 5, 3)],
}


def _int_taps(ssize: int, dsize: int):
    """per destination: (first source index, integer weights), and the common denominator."""
    tab = P.area_table(ssize, dsize)
    by = {}
    for d, s, w in tab:
        by.setdefault(d, []).append((s, float(w)))
    den = None
    for cand in range(1, 64):
        if all(abs(w * cand - round(w * cand)) < 1e-4 for v in by.values() for _, w in v):
            den = cand
            break
    assert den is not None
    out = []
    for d in range(dsize):
        v = by[d]
        assert [s for s, _ in v] == list(range(v[0][0], v[0][0] + len(v)))
        iw = [int(round(w * den)) for _, w in v]
        assert sum(iw) == den
        out.append((v[0][0], iw))
    return out, den


def test_chain_table_covers_every_tap_of_every_channel_once():
    for c, chain in CHAIN.items():
        offs = sorted(2 * e + (0 if st == "x" else 1) for st, e, _ in chain)
        assert offs == [3 * t + c for t in range(5)]
        for st, e, t in chain:
            assert 2 * e + (0 if st == "x" else 1) == 3 * t + c


@pytest.mark.parametrize("w,h", [(1920, 1080), (1280, 720)])
def test_vertical_first_integer_evaluation_equals_cv2_restatement(w, h):
    dw, dh = P.fit_size(w, h)
    assert (dw, dh) == (512, 288)
    xt, dx = _int_taps(w, dw)
    yt, dy = _int_taps(h, dh)
    D = dx * dy
    assert D % 2 == 1 and max(len(iw) for _, iw in xt) <= 5
    rng = np.random.default_rng(w)
    rows_out = 24                                    # a strip of area rows is enough (every row phase of the 3.75 period)
    r_lo = yt[40][0]
    r_hi = yt[40 + rows_out - 1][0] + len(yt[40 + rows_out - 1][1])
    frame = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    frame[r_lo:r_lo + 7, ::2] = 255                  # extremes exercise the rounding boundaries
    frame[r_lo + 7:r_lo + 11, ::3] = 0
    want = P.inter_area_resize(frame, dw, dh)[40:40 + rows_out]

    cols = list(range(140, 372))                     # the columns the 224-crop keeps, roughly
    xb0 = xt[cols[0]][0] * 3
    xb0 -= xb0 % 16                                  # stage rows start 16-byte aligned, like the bulk copies
    seg = (xt[cols[-1]][0] + 5) * 3 - xb0 + 16
    seg = (seg + 15) // 16 * 16
    got = np.zeros((rows_out, len(cols), 3), np.uint8)
    for r in range(rows_out):
        sy0, iyw = yt[40 + r]
        # vertical first: 16-bit lanes of (iy * byte), summed over the taps of this output row
        V = np.zeros(seg + 64, np.int64)
        for j, iy in enumerate(iyw):
            V[:seg] += iy * frame[sy0 + j].reshape(-1)[xb0:xb0 + seg].astype(np.int64)
        assert V.max() <= 255 * dy < 65536           # a lane never overflows
        L, H = V[0::2], V[1::2]                      # even-byte stream, odd-byte stream (16-bit elements)
        for ci, dxo in enumerate(cols):
            sx0, iw = xt[dxo]
            iw = iw + [0] * (5 - len(iw))
            s0 = sx0 * 3 - xb0
            parity, ex = s0 & 1, s0 >> 1
            ey = ex + parity
            X = (H if parity else L)[ex:ex + 8]      # X_i = V[s0 + 2i]
            Y = (L if parity else H)[ey:ey + 8]      # Y_i = V[s0 + 1 + 2i]
            assert all(X[i] == V[s0 + 2 * i] for i in range(8)) and all(Y[i] == V[s0 + 1 + 2 * i] for i in range(7))
            for c in range(3):
                n = sum(int(iw[t]) * int((X if st == "x" else Y)[e]) for st, e, t in CHAIN[c])
                got[r, ci, c] = (2 * n + D) // (2 * D)
    assert np.array_equal(got, want[:, cols])
