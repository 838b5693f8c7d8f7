"""Executable design model of csrc/octree.cu: the quadtree distribution on the points SORTED by their full-depth path
code, where a node is a run of the sorted array and the whole state is one depth per point.  Checked here against the
oracle's explicit-node version (oracle/svo_octree_oracle.c) on random point sets:  python tools/model_octree.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
MAXD = 12


def path_code(xr, yr, width, height, n_ini, hx):
    """root << 24 | 12 two-bit digits (digit d at bits 23-2d..22-2d); boxes halve with ceil like DivideNode."""
    r = int(np.float32(xr) / hx)
    r = min(max(r, 0), n_ini - 1)
    ulx, urx = int(hx * np.float32(r)), int(hx * np.float32(r + 1))
    uly, bry = 0, height
    code = r << 24
    for d in range(MAXD):
        sx = ulx + (urx - ulx + 1) // 2
        sy = uly + (bry - uly + 1) // 2
        q = 0
        if xr < sx: urx = sx
        else: ulx = sx; q |= 1
        if yr < sy: bry = sy
        else: uly = sy; q |= 2
        code |= q << (22 - 2 * d)
    return code


def shared_levels(a, b):
    x = a ^ b
    if x >> 24: return -1
    if x == 0: return MAXD
    return (23 - (x.bit_length() - 1)) >> 1


def distribute(xs, ys, score, rect, N):
    x0, y0, x1, y1 = rect
    n = len(xs)
    width, height = x1 - x0, y1 - y0
    if n == 0: return []
    # roundf: half away from zero
    n_ini = int(np.floor(np.float32(width) / np.float32(height) + np.float32(0.5)))
    n_ini = min(max(n_ini, 1), 16)
    hx = np.float32(width) / np.float32(n_ini)
    code = [path_code(int(xs[i]) - x0, int(ys[i]) - y0, width, height, n_ini, hx) for i in range(n)]
    packed = [int(xs[i]) | int(ys[i]) << 12 | int(score[i]) << 24 for i in range(n)]
    order = sorted(range(n), key=lambda i: (code[i], packed[i]))
    c = [code[i] for i in order]
    sh = [-1] + [shared_levels(c[i - 1], c[i]) for i in range(1, n)]
    dep = [0] * n

    def start(i, d): return i == 0 or sh[i] < d[i]
    def last(i, d): return i + 1 == n or start(i + 1, d)

    nn = sum(start(i, dep) for i in range(n))
    finish = False
    while not finish:
        prev = nn
        dep = [dep[i] + (0 if (start(i, dep) and last(i, dep)) or dep[i] >= MAXD else 1) for i in range(n)]
        nn = sum(start(i, dep) for i in range(n))
        ne = sum(start(i, dep) and not last(i, dep) and dep[i] < MAXD for i in range(n))
        if nn >= N or nn == prev:
            finish = True
        elif nn + 3 * ne > N:
            while not finish:
                prev = nn
                ns = [i for i in range(n) if start(i, dep)] + [n]
                cand = []
                for k in range(nn):
                    cnt = ns[k + 1] - ns[k]
                    d = dep[ns[k]]
                    if cnt > 1 and d < MAXD:
                        gain = sum(sh[i] == d for i in range(ns[k] + 1, ns[k + 1]))
                        cand.append((-cnt, k, gain))
                cand.sort()
                tot = nn
                for (_, k, gain) in cand:
                    for i in range(ns[k], ns[k + 1]): dep[i] += 1
                    tot += gain
                    if tot >= N: break
                nn = tot
                if nn >= N or nn == prev: finish = True
    ns = [i for i in range(n) if start(i, dep)] + [n]
    out = []
    for k in range(len(ns) - 1):
        best = max(range(ns[k], ns[k + 1]), key=lambda i: (packed[order[i]] >> 24, -(packed[order[i]] & 0xffffff)))
        out.append(order[best])
    return out


def main():
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    for trial in range(200):
        w = int(rng.integers(40, 1300)); h = int(rng.integers(40, 420))
        x0, y0 = 31, 31
        n = int(rng.integers(1, 2500))
        mode = trial % 4
        if mode == 0:
            px = rng.integers(0, w, n); py = rng.integers(0, h, n)
        elif mode == 1:    # clustered
            cx = rng.integers(0, w, 6); cy = rng.integers(0, h, 6)
            k = rng.integers(0, 6, n)
            px = np.clip(cx[k] + rng.normal(0, 12, n), 0, w - 1).astype(int); py = np.clip(cy[k] + rng.normal(0, 12, n), 0, h - 1).astype(int)
        elif mode == 2:    # a dense block
            px = rng.integers(0, min(w, 40), n); py = rng.integers(0, min(h, 40), n)
        else:              # one row / one column
            px = rng.integers(0, w, n); py = np.full(n, h // 2)
        pts = np.unique(np.stack([py, px], 1), axis=0)     # unique, raster order
        ys = pts[:, 0] + y0; xs = pts[:, 1] + x0
        sc = rng.integers(1, 5 if trial % 3 == 0 else 200, len(xs))
        N = int(rng.integers(1, 600))
        rect = (x0, y0, x0 + w, y0 + h)
        a = list(O.distribute_octree(xs, ys, sc, rect, N))
        b = distribute(xs, ys, sc, rect, N)
        assert a == b, (trial, w, h, len(xs), N, a[:10], b[:10])
    print("model_octree: sorted-run model == explicit-node oracle on 200 random point sets")


if __name__ == "__main__":
    main()
