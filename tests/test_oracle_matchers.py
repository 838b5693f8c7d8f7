"""Oracle matchers vs literal pure-Python transcriptions of the reference loops
(src/pnpmatch.cc:14-30, :75-95/:99-153 pass 1, :173-197 pass 2) on small cases."""
import ctypes as C

import numpy as np
import pytest

from oracle import oracle as O


def ham(a, b):
    return int(np.unpackbits(a ^ b).sum())


def ref_greedy(rows, cur, mode, claimed):
    """Transcription of the inner/outer loops; returns per-row (idx, best, second, took)."""
    claimed = claimed.copy()
    out = []
    for i in range(len(rows)):
        best, second, idx = 256, 256, -1
        for j in range(len(cur)):
            if claimed[j]:
                continue
            d = ham(rows[i], cur[j])
            if d < best:
                second, best, idx = best, d, j
        if mode == 0:
            take = best < 15
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                take = best < 30 and np.float32(second) / np.float32(best) > 2
        take = bool(take) and idx >= 0
        if take:
            claimed[idx] = 1
        out.append((idx, best, second, take))
    return out, claimed


def make_case(rng, M, N, flip_ps=(0.0, 0.01, 0.03, 0.08, 0.2)):
    cur = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    rows = np.empty((M, 32), np.uint8)
    for i in range(M):
        src = cur[rng.integers(0, N)]
        p = flip_ps[rng.integers(0, len(flip_ps))]
        flips = np.packbits(rng.random(256) < p, bitorder="little")
        rows[i] = src ^ flips
    return rows, cur


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_greedy_matches_transcription(mode, seed):
    rng = np.random.default_rng(seed)
    rows, cur = make_case(rng, 70, 50)
    cur[10] = cur[11]                      # exact duplicate columns: first-min tie-break
    claimed0 = (rng.random(50) < 0.1).astype(np.uint8)
    ref, claimed_ref = ref_greedy(rows, cur, mode, claimed0)
    got = O.match_greedy(rows, cur, mode, claimed=claimed0)
    assert [r[0] for r in ref] == list(got["best_idx"])
    assert [r[1] for r in ref] == list(got["best"])
    assert [r[2] for r in ref] == list(got["second"])
    assert [int(r[3]) for r in ref] == list(got["row_claimed"])
    assert (claimed_ref == got["claimed"]).all()


def test_hamming_swar_equals_popcount():
    rng = np.random.default_rng(3)
    a = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    b = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    for x, y in zip(a, b):
        assert O.hamming(x, y) == ham(x, y)
        assert O.lib().svo_o_hamming_popcnt(x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p)) == ham(x, y)
    assert O.hamming(a[0], a[0]) == 0
    assert O.hamming(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_greedy_empty_and_all_claimed():
    rng = np.random.default_rng(4)
    rows, cur = make_case(rng, 5, 8)
    got = O.match_greedy(rows, cur, 1, claimed=np.ones(8, np.uint8))
    assert (got["best_idx"] == -1).all() and (got["best"] == 256).all() and not got["row_claimed"].any()
    got = O.match_greedy(rows[:0], cur, 0)
    assert len(got["best_idx"]) == 0


def test_disp2depth():
    d = np.array([0.0, -1.0, 2.0, 0.5], np.float32)
    z = O.disp2depth(d, 379.8145)
    assert z[0] == -1.0 and z[1] == np.float32(379.8145) / np.float32(-1.0)
    assert z[2] == np.float32(379.8145) / np.float32(2.0)


def test_pass2_over_the_free_columns_only_is_the_same_scan():
    """The invariant behind k_free_cols: a column claimed before pass 2 starts is skipped by
    every pass-2 row (src/pnpmatch.cc:176), so scanning only the free columns — in ascending order, carrying the
    original index — gives the same claims, the same claim rows and the same (best, second) as scanning all of them."""
    rng = np.random.default_rng(11)
    for trial in range(6):
        M, N = int(rng.integers(50, 400)), int(rng.integers(60, 500))
        rows, cur = make_case(rng, M, N)
        claimed = (rng.random(N) < (0.0, 0.35, 0.8, 1.0, 0.5, 0.1)[trial]).astype(np.uint8)
        claim_row = np.where(claimed, rng.integers(0, 1000, N), -1).astype(np.int32)
        full = O.match_greedy(rows, cur, 1, claimed=claimed.copy(), claim_row=claim_row.copy(), row_base=1000)
        free = np.nonzero(claimed == 0)[0]
        sub = O.match_greedy(rows, cur[free], 1, row_base=1000)
        assert (full["row_claimed"] == sub["row_claimed"]).all()
        # claims of the compacted scan, mapped back to original column indices
        back = claim_row.copy()
        took = sub["claim_row"] >= 0
        back[free[took]] = sub["claim_row"][took]
        assert (full["claim_row"] == back).all()
        idx = sub["best_idx"].copy()
        ok = idx >= 0
        idx[ok] = free[idx[ok]]
        assert (full["best_idx"] == idx).all() and (full["best"] == sub["best"]).all() and (full["second"] == sub["second"]).all()


def test_project_map_formula_and_edge_cases():
    """The oracle's projection windows (opt-in pass-2 mode) against a plain numpy float32 restatement."""
    rng = np.random.default_rng(5)
    T = np.eye(4, dtype=np.float32); T[:3, 3] = [0.1, 0.0, 0.5]; T[0, 1] = np.float32(0.01); T[1, 0] = np.float32(-0.01)
    K4 = (np.float32(707.0912), np.float32(707.0912), np.float32(601.8873), np.float32(183.1104))
    xyz = np.stack([rng.uniform(-20, 20, 500), rng.uniform(-5, 5, 500), rng.uniform(-2, 60, 500)], 1).astype(np.float32)
    octv = rng.integers(0, 8, 500).astype(np.int32)
    out = O.project_map(xyz, octv, T, K4, 1241, 376, th=7.0)
    _, _, ls, _ = O.geometry(1241, 376, 8, 1.2, 500)
    for i in range(500):
        X, Y, Z = xyz[i]
        xc = ((T[0, 0] * X + T[0, 1] * Y) + T[0, 2] * Z) + T[0, 3]
        yc = ((T[1, 0] * X + T[1, 1] * Y) + T[1, 2] * Z) + T[1, 3]
        zc = ((T[2, 0] * X + T[2, 1] * Y) + T[2, 2] * Z) + T[2, 3]
        exp = (np.float32(0), np.float32(0), np.float32(-1))
        if zc > 0:
            iz = np.float32(1) / zc
            u = (K4[0] * xc) * iz + K4[2]; v = (K4[1] * yc) * iz + K4[3]
            if 0 <= u < 1241 and 0 <= v < 376:
                exp = (u, v, np.float32(7.0) * ls[octv[i]])
        assert tuple(out[i]) == tuple(np.float32(e) for e in exp), i
    assert (out[:, 2] > 0).sum() > 50 and (out[:, 2] < 0).sum() > 50
