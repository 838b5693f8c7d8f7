#!/usr/bin/env python3
"""Data-parallel model of libstdc++ introselect + partition (retainBest replay).

This is the design model for the CUDA select kernel (csrc/select.cu): every
step below is a map / prefix-sum / gather, i.e. what one thread block does
between barriers.  It is checked here against the real std::nth_element /
std::partition through the oracle shim.  Development aid; not on any product path.
"""
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
INF = 1 << 60


def hoare_step(key, val, lo, hi, pivot):
    """__unguarded_partition(first=lo, last=hi, pivot) with comp = greater; returns cut."""
    seg = key[lo:hi]
    is_l = ~(seg > pivot)          # left scan stops where !(x > pivot)
    is_r = ~(pivot > seg)          # right scan stops where !(pivot > x)
    lpos = lo + np.nonzero(is_l)[0]
    rpos = lo + np.nonzero(is_r)[0][::-1]
    kmax = min(len(lpos), len(rpos))
    m = int(np.count_nonzero(lpos[:kmax] < rpos[:kmax]))
    a, b = lpos[:m], rpos[:m]
    key[a], key[b] = key[b].copy(), key[a].copy()
    val[a], val[b] = val[b].copy(), val[a].copy()
    c1 = lpos[m] if m < len(lpos) else INF
    c2 = rpos[m - 1] if m >= 1 else INF
    return int(min(c1, c2))


def swap(key, val, i, j):
    key[i], key[j] = key[j], key[i]
    val[i], val[j] = val[j], val[i]


def move_median_to_first(key, val, r, a, b, c):
    gt = lambda i, j: key[i] > key[j]
    if gt(a, b):
        if gt(b, c): swap(key, val, r, b)
        elif gt(a, c): swap(key, val, r, c)
        else: swap(key, val, r, a)
    elif gt(a, c): swap(key, val, r, a)
    elif gt(b, c): swap(key, val, r, c)
    else: swap(key, val, r, b)


def adjust_heap(key, val, first, hole, length, vk, vv):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if key[first + child] > key[first + child - 1]:
            child -= 1
        key[first + hole] = key[first + child]; val[first + hole] = val[first + child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        key[first + hole] = key[first + child - 1]; val[first + hole] = val[first + child - 1]
        hole = child - 1
    parent = (hole - 1) // 2
    while hole > top and key[first + parent] > vk:
        key[first + hole] = key[first + parent]; val[first + hole] = val[first + parent]
        hole = parent
        parent = (hole - 1) // 2
    key[first + hole] = vk; val[first + hole] = vv


def heap_select(key, val, first, middle, last):
    length = middle - first
    if length >= 2:
        parent = (length - 2) // 2
        while True:
            adjust_heap(key, val, first, parent, length, key[first + parent], val[first + parent])
            if parent == 0:
                break
            parent -= 1
    for i in range(middle, last):
        if key[i] > key[first]:
            vk, vv = key[i], val[i]
            key[i] = key[first]; val[i] = val[first]
            adjust_heap(key, val, first, 0, length, vk, vv)


def introselect(key, val, nth, depth_limit=None):
    n = len(key)
    first, last = 0, n
    if n == 0 or nth >= n:
        return
    if depth_limit is None:
        depth_limit = 2 * (n.bit_length() - 1)
    while last - first > 3:
        if depth_limit == 0:
            heap_select(key, val, first, nth + 1, last)
            swap(key, val, first, nth)
            return
        depth_limit -= 1
        mid = first + (last - first) // 2
        move_median_to_first(key, val, first, first + 1, mid, last - 1)
        cut = hoare_step(key, val, first + 1, last, key[first])
        if cut <= nth:
            first = cut
        else:
            last = cut
    # insertion sort of <= 3 elements, comp = greater (stable shift)
    for i in range(first + 1, last):
        k, v = key[i], val[i]
        j = i
        if k > key[first]:
            while j > first:
                key[j] = key[j - 1]; val[j] = val[j - 1]; j -= 1
        else:
            while k > key[j - 1]:
                key[j] = key[j - 1]; val[j] = val[j - 1]; j -= 1
        key[j] = k; val[j] = v


def partition_ge(key, val, lo, hi, amb):
    """std::partition(lo, hi, x >= amb) (bidirectional version); returns the partition point."""
    seg = key[lo:hi] >= amb
    fpos = lo + np.nonzero(~seg)[0]
    tpos = lo + np.nonzero(seg)[0][::-1]
    kmax = min(len(fpos), len(tpos))
    m = int(np.count_nonzero(fpos[:kmax] < tpos[:kmax]))
    a, b = fpos[:m], tpos[:m]
    key[a], key[b] = key[b].copy(), key[a].copy()
    val[a], val[b] = val[b].copy(), val[a].copy()
    return lo + int(seg.sum())


def retain_best(resp, n_points, depth_limit=None):
    key = np.array(resp, np.float32)
    val = np.arange(len(key), dtype=np.int32)
    n = len(key)
    if n_points < 0 or n <= n_points:
        return val, key
    if n_points == 0:
        return val[:0], key[:0]
    introselect(key, val, n_points - 1, depth_limit)
    amb = key[n_points - 1]
    end = partition_ge(key, val, n_points, n, amb)
    return val[:end], key[:end]


def _check():
    import ctypes as C
    from oracle import oracle as O
    rng = np.random.default_rng(0)
    bad = 0
    for trial in range(3000):
        n = int(rng.integers(1, 400)) if trial % 3 else int(rng.integers(400, 20000))
        kind = trial % 4
        if kind == 0: resp = rng.integers(20, 60, n).astype(np.float32)            # heavy ties (FAST scores)
        elif kind == 1: resp = rng.normal(0, 1, n).astype(np.float32)              # Harris-like
        elif kind == 2: resp = np.sort(rng.integers(0, 1000, n)).astype(np.float32)
        else: resp = rng.integers(0, 3, n).astype(np.float32)
        npts = int(rng.integers(0, n + 2))
        i0, r0 = O.retain_best(resp, npts)
        i1, r1 = retain_best(resp, npts)
        if len(i0) != len(i1) or (i0 != i1).any():
            bad += 1; print("MISMATCH retain", trial, n, npts)
        # forced depth limits -> heap-select fallback against std::__introselect
        if n > 3:
            nth = int(rng.integers(0, n))
            for dl in (0, 1, 2):
                k0 = resp.copy(); v0 = np.arange(n, dtype=np.int32)
                O.lib().svo_o_introselect(k0.ctypes.data_as(C.c_void_p), v0.ctypes.data_as(C.c_void_p), n, nth, dl)
                k1 = resp.copy(); v1 = np.arange(n, dtype=np.int32)
                introselect(k1, v1, nth, dl)
                if (v0 != v1).any():
                    bad += 1; print("MISMATCH introselect", trial, n, nth, dl)
    print("done, mismatches:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if _check() else 0)
