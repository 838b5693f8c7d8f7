"""Executable models of two index / algebra arguments the round-2 kernels rest on (the kernels themselves are checked on
the GPU; these run everywhere and sweep cases the GPU tests do not enumerate).

* k_blur<true> (csrc/describe.cu): the staged tile row with its reflect-101 margins.  For EVERY level width from 8 to
  2600 and every tile of the row, every tap of every valid output pixel must read the reflect-101 pixel — including widths
  that are multiples of 16 (no padding byte to reuse), rows that end 1-2 pixels into their last tile, and single-tile rows.
* k_harris4 (csrc/describe.cu): the regrouped Sobel sums (column sums s, row differences v) equal the 3x3 form of
  k_harris / cv::ORB's HarrisResponses on random patches.
"""
import numpy as np

BLUR_RB = 4 * 128 + 32


def reflect101(i, n):
    if i < 0:
        i = -i
    if i >= n:
        i = 2 * (n - 1) - i
    return i


def blur_geometry(Lw):
    """svo_api.cu:build_geometry (blur tiles, whole-warp form)."""
    sp = (Lw + 15) // 16 * 16
    quads = (Lw + 3) // 4
    tiles_x = (quads + 127) // 128
    tq = ((quads + tiles_x - 1) // tiles_x + 31) // 32 * 32
    return sp, tiles_x, tq


def staged_row(pix, Lw, sp, tq, tx):
    """One staged row of tile tx as k_blur<true> builds it: TMA copy of [xs, xe) at offset xs - xv, then the margins."""
    xa = 4 * tq * tx
    xs, xe = max(xa - 16, 0), min(xa + 4 * tq + 16, sp)
    xv = xa - 16
    row = np.full(BLUR_RB, -1, np.int64)                      # -1: never written (shared memory garbage)
    glob = np.zeros(sp, np.int64); glob[:Lw] = pix            # the pitched row in HBM: padding bytes are 0
    assert xe > xs and (xe - xs) % 16 == 0 and (xs - xv) + (xe - xs) <= BLUR_RB
    row[xs - xv: xs - xv + (xe - xs)] = glob[xs:xe]
    if xa == 0:
        row[15], row[14], row[13] = row[17], row[18], row[19]
    if Lw > xs + 4 and Lw <= xa + 4 * tq + 8:
        e = Lw - xv
        assert e + 2 < BLUR_RB
        row[e], row[e + 1], row[e + 2] = row[e - 2], row[e - 3], row[e - 4]
    return row, xa, xv


def test_blur_margins_give_reflect101_for_every_width():
    rng = np.random.default_rng(0)
    checked = 0
    for Lw in list(range(8, 700)) + list(range(1020, 1060)) + list(range(1230, 1260)) + list(range(2540, 2600)):
        sp, tiles_x, tq = blur_geometry(Lw)
        assert tq <= 128 and tq * tiles_x * 4 >= Lw and (tiles_x - 1) * tq * 4 < Lw        # every tile holds a pixel
        pix = rng.integers(1, 256, Lw)
        for tx in range(tiles_x):
            row, xa, xv = staged_row(pix, Lw, sp, tq, tx)
            for tid in range(tq):
                x0 = xa + 4 * tid
                if x0 >= Lw:
                    continue
                lx = x0 - xv
                assert lx - 4 >= 0 and lx + 8 <= BLUR_RB                                 # the three word loads stay inside the row
                for k in range(4):
                    if x0 + k >= Lw:
                        continue
                    for tap in range(-3, 4):
                        p = x0 + k + tap
                        got = row[p - xv]
                        assert got == pix[reflect101(p, Lw)], (Lw, tx, tid, k, tap)
                        checked += 1
    assert checked > 1_000_000


def test_harris_regrouping_is_the_sobel_form():
    rng = np.random.default_rng(1)
    for _ in range(300):
        P = rng.integers(0, 256, (9, 9)).astype(np.int64)       # patch rows y-4..y+4, columns x-4..x+4
        a = b = c = 0
        for dy in range(7):                                     # k_harris / HarrisResponses: 3x3 Sobel pair per block position
            for dx in range(7):
                q = P[dy:dy + 3, dx:dx + 3]
                Ix = (q[1, 2] - q[1, 0]) * 2 + (q[0, 2] - q[0, 0]) + (q[2, 2] - q[2, 0])
                Iy = (q[2, 1] - q[0, 1]) * 2 + (q[2, 0] - q[0, 0]) + (q[2, 2] - q[0, 2])
                a += Ix * Ix; b += Iy * Iy; c += Ix * Iy
        a2 = b2 = c2 = 0
        for r in range(7):                                      # k_harris4: lane r, rows r..r+2
            p0, p1, p2 = P[r], P[r + 1], P[r + 2]
            s = p0 + 2 * p1 + p2
            v = p2 - p0
            for dx in range(7):
                Ix = s[dx + 2] - s[dx]
                Iy = v[dx] + 2 * v[dx + 1] + v[dx + 2]
                a2 += Ix * Ix; b2 += Iy * Iy; c2 += Ix * Iy
        assert (a, b, c) == (a2, b2, c2)
