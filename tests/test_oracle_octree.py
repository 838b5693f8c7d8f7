"""The opt-in quadtree keypoint distribution (oracle/svo_octree_oracle.c — the definition of that mode; the reference has
no such stage, so nothing here is pinned to reference output).  The explicit-node C version is checked against the
independent sorted-run model the device kernel follows (tools/model_octree.py) and against the properties that make it
a distribution: one keypoint per final node, the count rule of DistributeOctTree, spatial coverage."""
import os
import sys

import numpy as np
import pytest

import synth
from oracle import oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import model_octree as M  # noqa: E402


def random_points(rng, w, h, n, mode):
    if mode == 0:
        px, py = rng.integers(0, w, n), rng.integers(0, h, n)
    elif mode == 1:   # clusters
        c = rng.integers(0, 6, n)
        cx, cy = rng.integers(0, w, 6), rng.integers(0, h, 6)
        px = np.clip(cx[c] + rng.normal(0, 12, n), 0, w - 1).astype(int)
        py = np.clip(cy[c] + rng.normal(0, 12, n), 0, h - 1).astype(int)
    elif mode == 2:   # one dense block
        px, py = rng.integers(0, min(w, 40), n), rng.integers(0, min(h, 40), n)
    else:             # a single row
        px, py = rng.integers(0, w, n), np.full(n, h // 2)
    pts = np.unique(np.stack([py, px], 1), axis=0)   # unique positions, raster order
    return pts[:, 1] + 31, pts[:, 0] + 31


@pytest.mark.parametrize("mode", [0, 1, 2, 3])
def test_explicit_nodes_equal_sorted_run_model(mode):
    rng = np.random.default_rng(100 + mode)
    for trial in range(25):
        w, h = int(rng.integers(40, 1300)), int(rng.integers(40, 420))
        xs, ys = random_points(rng, w, h, int(rng.integers(1, 1500)), mode)
        sc = rng.integers(1, 4 if trial % 2 else 200, len(xs))   # heavy score ties every other trial
        N = int(rng.integers(1, 500))
        rect = (31, 31, 31 + w, 31 + h)
        assert list(O.distribute_octree(xs, ys, sc, rect, N)) == M.distribute(xs, ys, sc, rect, N)


def test_counts_and_uniqueness():
    rng = np.random.default_rng(7)
    w, h = 1179, 314
    xs, ys = random_points(rng, w, h, 6000, 0)
    sc = rng.integers(1, 200, len(xs))
    rect = (31, 31, 31 + w, 31 + h)
    for N in (1, 50, 434, 1737):
        k = O.distribute_octree(xs, ys, sc, rect, N)
        assert len(set(k.tolist())) == len(k)
        # splitting stops at the first count >= N; one split adds at most 3 nodes; the first round always runs
        assert max(N, 1) <= len(k) <= max(N + 2, 16)
    # fewer points than N: every point is alone in its node in the end
    few = rng.choice(len(xs), 300, replace=False); few.sort()
    k = O.distribute_octree(xs[few], ys[few], sc[few], rect, 434)
    assert sorted(k.tolist()) == list(range(300))
    assert len(O.distribute_octree(xs[:0], ys[:0], sc[:0], rect, 10)) == 0
    assert list(O.distribute_octree(xs[:1], ys[:1], sc[:1], rect, 10)) == [0]


def test_spreads_keypoints_where_retain_best_clusters():
    """An image whose strong texture sits in one corner: retainBest spends the budget there, the quadtree covers the
    frame (the reason ORB-SLAM2-lineage front-ends use it)."""
    img = synth.texture((240, 400), 11).astype(np.float32)
    weak = 128 + (img - 128) * 0.35
    weak[:100, :150] = img[:100, :150]
    img = np.clip(np.rint(weak), 0, 255).astype(np.uint8)
    k0, d0, _ = O.orb(img, 300)
    k1, d1, _ = O.orb(img, 300, distribution=1)
    assert len(k1) > 0 and d1.shape == (len(k1), 32)

    def cells(k):
        lv0 = k[k["octave"] == 0]
        return len(set(zip((lv0["x"] // 40).astype(int).tolist(), (lv0["y"] // 40).astype(int).tolist())))
    assert cells(k1) > cells(k0)
    # per-level counts follow the quota rule
    lw, lh, ls, quota = O.geometry(400, 240, 8, 1.2, 300)
    for l in range(8):
        assert (k1["octave"] == l).sum() <= max(quota[l] + 2, 16)


def test_octree_mode_shares_everything_else_with_the_parity_path():
    """A keypoint both modes select gets the same angle, response and descriptor: only the selection differs."""
    img = synth.texture((240, 400), 12)
    k0, d0, _ = O.orb(img, 500)
    k1, d1, _ = O.orb(img, 500, distribution=1)
    key0 = {(float(k["x"]), float(k["y"]), int(k["octave"])): i for i, k in enumerate(k0)}
    common = 0
    for j, k in enumerate(k1):
        i = key0.get((float(k["x"]), float(k["y"]), int(k["octave"])))
        if i is None:
            continue
        common += 1
        assert k0[i]["angle"] == k["angle"] and k0[i]["response"] == k["response"] and (d0[i] == d1[j]).all()
    assert common > 50
