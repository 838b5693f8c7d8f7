"""GPU parity: every stage of libsvo_b200.so (through the C ABI) against the CPU oracle and the
committed OpenCV golden vectors.  Integer/byte/index work must be bit-exact; floats (angle,
response, keypoint coordinates) are compared by bit pattern too; sparse-stereo sub-pixel
disparity/depth within 1e-3 (north_star's tolerance)."""
import glob
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.fixture(scope="module")
def svo():
    import svo as S
    return S


@pytest.fixture(scope="module")
def ctxK(svo):
    c = svo.Context(1241, 376, nfeatures=2000, max_batch=4, lanes=2, max_rows=5000)
    yield c
    c.close()


def assert_kp_equal(kp, ref):
    assert len(kp) == len(ref)
    for f in ("x", "y", "size", "angle", "response"):
        bad = np.nonzero(bits(kp[f]) != bits(ref[f]))[0]
        assert len(bad) == 0, "%s differs at %s (%d of %d)" % (f, bad[:5], len(bad), len(ref))
    assert (kp["octave"] == ref["octave"]).all()


def test_geometry(ctxK):
    lw, lh, ls, q = ctxK.geometry()
    olw, olh, ols, oq = O.geometry(1241, 376, 8, 1.2, 2000)
    assert (lw == olw).all() and (lh == olh).all() and (q == oq).all()
    assert (bits(ls) == bits(ols)).all()


def test_stages_vs_oracle(ctxK, svo):
    """pyramid, blur, FAST list, first and second cull per level (order included)."""
    img = synth.texture(synth.K_SHAPE, 31)
    ctxK.extract(img, cam=0)
    kp, desc, pyr = O.orb(img, 2000, with_pyramid=True)
    lw, lh, ls, quota = O.geometry(1241, 376, 8, 1.2, 2000)
    for l in range(8):
        lev = pyr.level(l)
        assert (ctxK.tap_image(0, l) == lev).all(), "pyramid level %d" % l
        assert (ctxK.tap_image(0, l, blurred=True) == pyr.level(l, True)).all(), "blur level %d" % l
        xs, ys, sc = O.fast_nms(lev, 20, 31)
        f = ctxK.tap_list(0, svo.TAP_FAST, l)
        assert len(f) == len(xs), "FAST count level %d: %d vs %d" % (l, len(f), len(xs))
        assert (f[:, 0] == xs).all() and (f[:, 1] == ys).all() and (f[:, 2] == sc).all(), "FAST list level %d" % l
        idx, _ = O.retain_best(sc.astype(np.float32), 2 * int(quota[l]))
        s1 = ctxK.tap_list(0, svo.TAP_SELECT1, l)
        assert len(s1) == len(idx), "first cull count level %d" % l
        assert (s1[:, 0] == xs[idx]).all() and (s1[:, 1] == ys[idx]).all(), "first cull order level %d" % l
        hr = O.harris(lev, xs[idx], ys[idx])
        idx2, r2 = O.retain_best(hr, int(quota[l]))
        s2 = ctxK.tap_list(0, svo.TAP_SELECT2, l)
        assert len(s2) == len(idx2), "second cull count level %d" % l
        assert (s2[:, 0] == xs[idx][idx2]).all() and (s2[:, 1] == ys[idx][idx2]).all(), "second cull order level %d" % l
        assert (s2[:, 2].view(np.uint32) == bits(r2)).all(), "harris bits level %d" % l
    O.pyramid_free(pyr)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(G, "k2000*.npz"))), ids=lambda p: os.path.basename(p)[:-4])
def test_extract_vs_cv2_golden(ctxK, path):
    g = np.load(path)
    kp, desc = ctxK.extract(g["image"], cam=0)
    assert_kp_equal(kp, g["kp"])
    assert (desc == g["desc"]).all()


@pytest.mark.parametrize("seed", [41, 42])
def test_extract_vs_oracle(ctxK, seed):
    img = synth.texture(synth.K_SHAPE, seed)
    kp, desc = ctxK.extract(img, cam=1)
    okp, odesc, _ = O.orb(img, 2000)
    assert_kp_equal(kp, okp)
    assert (desc == odesc).all()


def test_extract_small_golden_and_other_sizes(svo):
    g = np.load(os.path.join(G, "s500_seed21.npz"))
    c = svo.Context(400, 240, nfeatures=500, max_rows=1000)
    kp, desc = c.extract(g["image"])
    assert_kp_equal(kp, g["kp"])
    assert (desc == g["desc"]).all()
    c.close()
    # odd size, few features wanted, strided input
    img = synth.texture((203, 330), 5)
    c = svo.Context(317, 203, nfeatures=300, max_rows=1000)
    kp, desc = c.extract(img[:, :317])
    okp, odesc, _ = O.orb(np.ascontiguousarray(img[:, :317]), 300)
    assert_kp_equal(kp, okp)
    assert (desc == odesc).all()
    # featureless image: zero keypoints, no error
    kp, desc = c.extract(np.full((203, 317), 90, np.uint8))
    assert len(kp) == 0
    c.close()


def test_extract_highres_8000(svo):
    img = synth.texture(synth.H_SHAPE, 51)
    c = svo.Context(2560, 720, nfeatures=8000, max_rows=8000)
    kp, desc = c.extract(img)
    okp, odesc, _ = O.orb(img, 8000)
    assert_kp_equal(kp, okp)
    assert (desc == odesc).all()
    c.close()


def test_retain_best_replay_vs_libstdcxx(ctxK):
    """The on-device introselect/partition replay against std::nth_element/std::partition."""
    rng = np.random.default_rng(0)
    for trial in range(60):
        n = int(rng.integers(1, 600)) if trial % 3 else int(rng.integers(600, 40000))
        kind = trial % 4
        if kind == 0:
            resp = rng.integers(20, 60, n).astype(np.float32)
        elif kind == 1:
            resp = rng.normal(0, 1e-5, n).astype(np.float32)
        elif kind == 2:
            resp = np.sort(rng.integers(0, 1000, n)).astype(np.float32)
        else:
            resp = rng.integers(0, 3, n).astype(np.float32)
        npts = int(rng.integers(0, n + 2))
        ref, _ = O.retain_best(resp, npts)
        got = ctxK.retain_best(resp, npts)
        assert len(got) == len(ref) and (got == ref).all(), (trial, n, npts)


def test_retain_best_forced_heap_select(ctxK):
    import ctypes as C
    rng = np.random.default_rng(1)
    for trial in range(12):
        n = int(rng.integers(5, 3000))
        resp = rng.integers(0, 50, n).astype(np.float32)
        npts = int(rng.integers(1, n))
        for dl in (0, 1, 3):
            k = resp.copy(); v = np.arange(n, dtype=np.int32)
            O.lib().svo_o_introselect(k.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), n, npts - 1, dl)
            amb = k[npts - 1]
            got = ctxK.retain_best(resp, npts, depth_limit=dl)
            assert (got[:npts] == v[:npts]).all(), (trial, n, npts, dl)
            assert (resp[got[npts:]] >= amb).all()


def test_match_bf_golden_and_oracle(ctxK):
    g = np.load(os.path.join(G, "bfmatch.npz"))
    idx, dist, keep = ctxK.match_bf(g["q"], g["t"])
    assert (idx == g["train"]).all() and (dist.astype(np.float32) == g["dist"]).all()
    rng = np.random.default_rng(2)
    q = rng.integers(0, 256, (2100, 32), dtype=np.uint8); t = rng.integers(0, 256, (3001, 32), dtype=np.uint8)
    t[5:900] = q[100:995]; t[1500] = t[7]
    oi, od, ok = O.match_bf(q, t)
    idx, dist, keep = ctxK.match_bf(q, t)
    assert (idx == oi).all() and (dist == od).all() and (keep == ok).all()


def noisy_copies(rng, base, n, ps=(0.0, 0.01, 0.03, 0.08, 0.2, 0.5)):
    src = base[rng.integers(0, len(base), n)]
    p = np.asarray(ps)[rng.integers(0, len(ps), n)]
    flips = np.packbits(rng.random((n, 256)) < p[:, None], axis=1, bitorder="little")
    return src ^ flips


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("seed", [0, 1])
def test_match_greedy_vs_oracle(ctxK, mode, seed):
    rng = np.random.default_rng(seed)
    cur = rng.integers(0, 256, (1999, 32), dtype=np.uint8)
    cur[300:340] = cur[100]                      # 41 identical columns: short-list overflow path
    rows = noisy_copies(rng, cur, 3000)
    rows[10:60] = cur[100]
    claimed0 = (rng.random(len(cur)) < 0.05).astype(np.uint8)
    live = (rng.random(len(rows)) < 0.9).astype(np.uint8)
    ref = O.match_greedy(rows, cur, mode, claimed=claimed0, row_live=live, row_base=7)
    got = ctxK.match_greedy(rows, cur, mode, claimed=claimed0, row_live=live, row_base=7)
    for k in ("row_claimed", "claimed", "claim_row"):
        assert (ref[k] == got[k]).all(), k
    lv = live.astype(bool)
    for k in ("best_idx", "best", "second"):
        assert (ref[k][lv] == got[k][lv]).all(), k


@pytest.mark.parametrize("mode", [0, 1])
def test_match_greedy_long_chains_and_list_overflow(ctxK, mode):
    """200 identical columns and 260 identical rows: every such row's short list overflows (exhaustive
    re-scan path) and each claim depends on the previous one (the fixed-point resolver needs one sweep per
    link of the chain); a second family of 90 near-duplicates stays within the short-list capacity."""
    rng = np.random.default_rng(11 + mode)
    cur = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    cur[400:600] = cur[7]
    cur[700:790] = cur[9]
    rows = noisy_copies(rng, cur, 2000)
    rows[100:360] = cur[7]
    rows[500:560] = cur[9]
    rows[900:960] = cur[9] ^ np.uint8(1)
    ref = O.match_greedy(rows, cur, mode, row_base=3)
    got = ctxK.match_greedy(rows, cur, mode, row_base=3)
    for k in ("row_claimed", "claimed", "claim_row", "best_idx", "best", "second"):
        assert (ref[k] == got[k]).all(), k
    assert ref["row_claimed"][100:360].sum() >= 150


def test_match_greedy_window_and_edge_cases(ctxK):
    rng = np.random.default_rng(3)
    cur = rng.integers(0, 256, (500, 32), dtype=np.uint8)
    rows = noisy_copies(rng, cur, 700)
    cur_xy = rng.uniform(0, 1000, (500, 2)).astype(np.float32)
    win = np.concatenate([rng.uniform(0, 1000, (700, 2)), rng.uniform(50, 400, (700, 1))], 1).astype(np.float32)
    ref = O.match_greedy(rows, cur, 1, win_uvr=win, cur_xy=cur_xy)
    got = ctxK.match_greedy(rows, cur, 1, win_uvr=win, cur_xy=cur_xy)
    for k in ("row_claimed", "claimed", "claim_row", "best_idx", "best", "second"):
        assert (ref[k] == got[k]).all(), k
    # everything already claimed
    got = ctxK.match_greedy(rows, cur, 0, claimed=np.ones(500, np.uint8))
    assert not got["row_claimed"].any() and (got["best_idx"] == -1).all() and (got["best"] == 256).all()


def test_disp2depth(ctxK):
    rng = np.random.default_rng(4)
    d = rng.uniform(-1, 60, 376 * 1241).astype(np.float32)
    d[::7] = 0
    assert (bits(ctxK.disp2depth(d, 379.8145)) == bits(O.disp2depth(d, 379.8145))).all()


def stereo_oracle(L, R, nf, bf, b):
    kl, dl, pl = O.orb(L, nf, with_pyramid=True)
    kr, dr, pr = O.orb(R, nf, with_pyramid=True)
    ur, dep, mr, sad = O.stereo_sparse(kl, dl, pl, kr, dr, pr, bf, b)
    O.pyramid_free(pl); O.pyramid_free(pr)
    return kl, dl, kr, dr, ur, dep, mr, sad


def test_stereo_sparse_vs_oracle(ctxK):
    cal = synth.KITTI_04_12
    bf, b = cal["bf"], cal["bf"] / cal["fx"]
    L, R, _ = synth.stereo_pair(synth.K_SHAPE, seed=61)
    kl, dl, kr, dr, ur, dep, mr, sad = stereo_oracle(L, R, 2000, bf, b)
    ctxK.extract(L, cam=0); ctxK.extract(R, cam=1)
    gur, gdep, gmr, gsad = ctxK.stereo_sparse(bf, b)
    assert len(gur) == len(ur)
    valid = dep > 0
    assert valid.sum() > 300, "synthetic pair should yield stereo matches (%d)" % valid.sum()
    assert ((gdep > 0) == valid).all()
    assert (gmr[valid] == mr[valid]).all() and (gsad[valid] == sad[valid]).all()
    assert np.abs(gur[valid] - ur[valid]).max() <= 1e-3
    assert (np.abs(gdep[valid] - dep[valid]) <= 1e-3 * np.abs(dep[valid])).all()
    assert (gur[~valid] == -1).all() and (gdep[~valid] == -1).all()


def test_batch_pipeline_vs_oracle(ctxK):
    """Full per-frame front-end through svo_batch_submit on two lanes, checked against the oracle."""
    cal = synth.KITTI_04_12
    bf, b = np.float32(cal["bf"]), np.float32(cal["bf"] / cal["fx"])
    seq = synth.Sequence(seed=2)
    frames = [seq.frame(t) for t in range(4)]
    ora = [stereo_oracle(L, R, 2000, bf, b) for L, R in frames]
    rng = np.random.default_rng(9)
    jobs = []
    for t in range(1, 4):
        prev_desc = ora[t - 1][1]
        live = (ora[t - 1][5] > 0).astype(np.uint8)          # map points exist where stereo depth > 0
        mp = synth.local_map([o[1] for o in ora[:t]], rows=5000, seed=t)
        mpr = np.full(5000, -1, np.int32)
        take = rng.permutation(5000)[:600]
        src = rng.integers(0, len(prev_desc), 600)
        mp[take] = prev_desc[src]; mpr[take] = src            # map rows that are the last frame's own points
        jobs.append(dict(left=frames[t][0], right=frames[t][1], bf=float(bf), baseline=float(b),
                         prev_desc=prev_desc, prev_live=live, map_desc=mp, map_prev_row=mpr))
    ctxK.batch_submit(0, jobs[:2])
    ctxK.batch_submit(1, jobs[2:])
    ctxK.batch_wait(0); ctxK.batch_wait(1)
    res = [ctxK.batch_result(0, 0), ctxK.batch_result(0, 1), ctxK.batch_result(1, 0)]
    for t, (job, r) in enumerate(zip(jobs, res), start=1):
        kl, dl, kr, dr, ur, dep, mr, sad = ora[t]
        assert r["status"] == 0
        assert_kp_equal(r["kp_left"], kl); assert_kp_equal(r["kp_right"], kr)
        assert (r["desc_left"] == dl).all() and (r["desc_right"] == dr).all()
        valid = dep > 0
        assert ((r["depth"] > 0) == valid).all() and r["n_stereo"] == valid.sum()
        assert np.abs(r["u_right"][valid] - ur[valid]).max() <= 1e-3
        oi, od, ok = O.match_bf(dl, job["prev_desc"])
        assert (r["bf_idx"] == oi).all() and (r["bf_dist"] == od).all() and (r["bf_keep"] == ok).all()
        p1 = O.match_greedy(job["prev_desc"], dl, 0, row_live=job["prev_live"])
        lv = job["prev_live"].astype(bool)
        assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
        for k in ("best_idx", "best", "second"):
            assert (r["p1_" + k][lv] == p1[k][lv]).all(), k
        live2 = np.ones(5000, np.uint8)
        m = job["map_prev_row"] >= 0
        live2[m] = 1 - p1["row_claimed"][job["map_prev_row"][m]]
        p2 = O.match_greedy(job["map_desc"], dl, 1, claimed=p1["claimed"], claim_row=p1["claim_row"],
                            row_live=live2, row_base=len(job["prev_desc"]))
        assert (r["p2_row_claimed"] == p2["row_claimed"]).all()
        assert (r["claim_row"] == p2["claim_row"]).all()
        assert p1["row_claimed"].sum() > 50 and p2["row_claimed"].sum() > 50


def _same_result(a, b):
    for k in a:
        if isinstance(a[k], np.ndarray):
            assert a[k].shape == b[k].shape and a[k].tobytes() == b[k].tobytes(), k
        else:
            assert a[k] == b[k], k


def test_batch_inputs_any_residency_and_layout(ctxK):
    """svo_batch_submit reads device buffers in place and gathers host buffers run by run: the results must not
    depend on where the inputs live or how they are laid out (device pointers, one contiguous host block shared by
    consecutive frames, strided image views, descriptor blocks at addresses that are not 16-byte aligned)."""
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    seq = synth.Sequence(seed=5)
    H, W = synth.K_SHAPE
    frames = [seq.frame(t) for t in range(3)]
    prevs = [ctxK.extract(frames[t][0], cam=0)[1] for t in range(3)]
    rng = np.random.default_rng(12)
    jobs = []
    for t in (1, 2):
        prev = prevs[t - 1]
        mp = synth.local_map(prevs[:t], rows=3000, seed=t)
        mpr = np.full(3000, -1, np.int32)
        take = rng.permutation(3000)[:400]; src = rng.integers(0, len(prev), 400)
        mp[take] = prev[src]; mpr[take] = src
        live = (rng.random(len(prev)) < 0.8).astype(np.uint8)
        jobs.append(dict(left=frames[t][0], right=frames[t][1], bf=bf, baseline=b, prev_desc=prev, prev_live=live,
                         map_desc=mp, map_prev_row=mpr))
    ctxK.batch_submit(0, jobs); ctxK.batch_wait(0)
    base = [ctxK.batch_result(0, i) for i in range(2)]
    assert base[0]["p1_row_claimed"].sum() > 20 and base[0]["p2_row_claimed"].sum() > 20

    # (1) everything device resident, passed as raw addresses
    dev = []
    for j in jobs:
        dev.append(dict(left=ctxK.to_device(j["left"]), right=ctxK.to_device(j["right"]), stride=W, bf=bf, baseline=b,
                        prev_desc=ctxK.to_device(j["prev_desc"]), n_prev=len(j["prev_desc"]),
                        prev_live=ctxK.to_device(j["prev_live"]),
                        map_desc=ctxK.to_device(j["map_desc"]), n_map=len(j["map_desc"]),
                        map_prev_row=ctxK.to_device(j["map_prev_row"])))
    ctxK.batch_submit(1, dev); ctxK.batch_wait(1)
    for i in range(2):
        _same_result(base[i], ctxK.batch_result(1, i))

    # (2) host inputs in awkward layouts: frame 0's left and right images adjacent in one block (one gathered run),
    # frame 1's images as strided views, descriptors at odd addresses, the same map block given to both frames
    block = np.zeros((2, H, W), np.uint8)
    wide = np.zeros((2, H, W + 37), np.uint8)
    raw = np.zeros(2 * (len(jobs[0]["prev_desc"]) + len(jobs[1]["prev_desc"])) * 32 + 64, np.uint8)
    odd, off = [], 4
    for i, j in enumerate(jobs):
        if i == 0:
            block[0] = j["left"]; block[1] = j["right"]; imgs = (block[0], block[1])
        else:
            wide[0, :, :W] = j["left"]; wide[1, :, :W] = j["right"]; imgs = (wide[0, :, :W], wide[1, :, :W])
        n = len(j["prev_desc"])
        v = raw[off:off + n * 32].reshape(n, 32); v[:] = j["prev_desc"]
        off += n * 32 + 4
        odd.append(dict(left=imgs[0], right=imgs[1], bf=bf, baseline=b, prev_desc=v, prev_live=j["prev_live"],
                        map_desc=jobs[0]["map_desc"], map_prev_row=None))
        assert v.ctypes.data % 16 != 0
    ref2 = [dict(j, map_desc=jobs[0]["map_desc"], map_prev_row=None) for j in jobs]
    ctxK.batch_submit(0, ref2); ctxK.batch_wait(0)
    want = [ctxK.batch_result(0, i) for i in range(2)]
    ctxK.batch_submit(1, odd); ctxK.batch_wait(1)
    for i in range(2):
        _same_result(want[i], ctxK.batch_result(1, i))


def test_batch_stage_combinations_and_fallback_path(ctxK):
    """Frames without a previous frame, without a map, with more previous rows than current keypoints can hold
    (the un-fused BF / pass-1 path), batches of different sizes on the same lane (each (n, stages) combination is
    its own captured graph) -- all against the oracle."""
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    seq = synth.Sequence(seed=8)
    frames = [seq.frame(t) for t in range(3)]
    descs = [ctxK.extract(f[0], cam=0)[1] for f in frames]
    rng = np.random.default_rng(5)
    big_prev = np.concatenate([descs[0], noisy_copies(rng, descs[1], 3000 - len(descs[0]))], 0)   # 3000 rows > kp_cap
    small_map = synth.local_map(descs[:2], rows=700, seed=3)

    def check(job, r):
        dl = r["desc_left"]
        assert r["status"] == 0
        if job.get("prev_desc") is not None:
            oi, od, ok = O.match_bf(dl, job["prev_desc"])
            assert (r["bf_idx"] == oi).all() and (r["bf_dist"] == od).all() and (r["bf_keep"] == ok).all()
            p1 = O.match_greedy(job["prev_desc"], dl, 0)
            assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
            for k in ("best_idx", "best", "second"):
                assert (r["p1_" + k] == p1[k]).all(), k
            claimed, claim_row, base = p1["claimed"], p1["claim_row"], len(job["prev_desc"])
        else:
            claimed, claim_row, base = None, None, 0
        if job.get("map_desc") is not None:
            p2 = O.match_greedy(job["map_desc"], dl, 1, claimed=claimed, claim_row=claim_row, row_base=base)
            assert (r["p2_row_claimed"] == p2["row_claimed"]).all()
            assert (r["claim_row"] == p2["claim_row"]).all()

    def img(t):
        return dict(left=frames[t][0], right=frames[t][1], bf=bf, baseline=b)

    batches = [
        [dict(img(1))],                                                               # extraction + stereo only
        [dict(img(1), prev_desc=descs[0]), dict(img(2), prev_desc=descs[1])],         # no map
        [dict(img(2), map_desc=small_map)],                                           # no previous frame
        [dict(img(2), prev_desc=big_prev, map_desc=small_map)],                       # n_prev > kp_cap: un-fused path
        [dict(img(1), prev_desc=descs[0], map_desc=small_map), dict(img(2), prev_desc=descs[1], map_desc=small_map),
         dict(img(2), prev_desc=descs[0], map_desc=small_map)],                       # three frames, fused path
        [dict(img(1), prev_desc=descs[0]), dict(img(2), prev_desc=descs[1])],         # replay of a cached graph
    ]
    for jobs in batches:
        ctxK.batch_submit(0, jobs); ctxK.batch_wait(0)
        for i, job in enumerate(jobs):
            check(job, ctxK.batch_result(0, i))
    # a featureless stereo pair next to a normal one: zero keypoints, nothing matched, nothing claimed, no error
    flat = np.full(synth.K_SHAPE, 117, np.uint8)
    jobs = [dict(left=flat, right=flat, bf=bf, baseline=b, prev_desc=descs[0], map_desc=small_map),
            dict(img(1), prev_desc=descs[0], map_desc=small_map)]
    ctxK.batch_submit(1, jobs); ctxK.batch_wait(1)
    r = ctxK.batch_result(1, 0)
    assert r["status"] == 0 and r["n_left"] == 0 and r["n_right"] == 0 and r["n_stereo"] == 0
    assert not r["p1_row_claimed"].any() and not r["p2_row_claimed"].any()
    assert (r["p1_best_idx"] == -1).all() and (r["p1_best"] == 256).all() and (r["p1_second"] == 256).all()
    check(jobs[1], ctxK.batch_result(1, 1))


# ---- colour input (the staging step before the path) -----------------------------
def test_extract_bgr_matches_oracle_and_gray_path(svo):
    H, W = 240, 400
    c = svo.Context(W, H, nfeatures=500, max_batch=2, lanes=1, max_rows=1000, max_channels=3)
    try:
        gray = synth.texture((H, W), 33)
        col = synth.colourise(gray, 1)
        g = O.bgr2gray(col)
        ref, rdesc, _ = O.orb(g, 500)
        kp, desc = c.extract(col)
        assert_kp_equal(kp, ref)
        assert (desc == rdesc).all()
        assert (c.tap_image(0, 0) == g).all()
        # strided colour rows (a padded cv::Mat), right camera slot
        pad = np.zeros((H, W + 7, 3), np.uint8); pad[:, :W] = col
        kp2, desc2 = c.extract(pad[:, :W], cam=1)
        assert_kp_equal(kp2, ref)
        assert (desc2 == rdesc).all()
        # the same context still takes gray input
        kp3, desc3 = c.extract(g)
        assert_kp_equal(kp3, ref)
        # batch: BGR host frames, BGR device frames and a gray frame in one call
        colr = synth.colourise(synth.texture((H, W), 34), 2)
        gr = O.bgr2gray(colr)
        rr, rrd, _ = O.orb(gr, 500)
        dl, dr = c.to_device(col), c.to_device(colr)
        c.batch_submit(0, [dict(left=col, right=colr, bf=100.0, baseline=0.5),
                           dict(left=dl, right=dr, stride=3 * W, channels=3, bf=100.0, baseline=0.5)])
        c.batch_wait(0)
        for i in range(2):
            r = c.batch_result(0, i)
            assert r["status"] == 0
            assert_kp_equal(r["kp_left"], ref); assert_kp_equal(r["kp_right"], rr)
            assert (r["desc_left"] == rdesc).all() and (r["desc_right"] == rrd).all()
        c.batch_submit(0, [dict(left=g, right=gr, bf=100.0, baseline=0.5)])
        c.batch_wait(0)
        r = c.batch_result(0, 0)
        assert_kp_equal(r["kp_left"], ref); assert_kp_equal(r["kp_right"], rr)
    finally:
        c.close()


def test_bgr_needs_a_colour_context(svo, ctxK):
    col = synth.colourise(synth.texture(synth.K_SHAPE, 3), 1)
    with pytest.raises(svo.SvoError) as e:
        ctxK.extract(col)
    assert e.value.code == svo.E_INVALID
    with pytest.raises(svo.SvoError) as e:
        ctxK.batch_submit(0, [dict(left=col, right=col, bf=100.0, baseline=0.5)])
    assert e.value.code == svo.E_INVALID


def test_context_for_a_small_image_leaves_a_larger_one_working(svo, ctxK):
    """Kernel attributes (k_fast's dynamic shared-memory limit) are per kernel, not per context: creating a context
    for a smaller geometry must not lower them under a context that is already serving 1241x376."""
    img = synth.texture(synth.K_SHAPE, 5)
    kp0, d0 = ctxK.extract(img)
    small = svo.Context(400, 240, nfeatures=300, max_batch=1, lanes=1, max_rows=500)
    try:
        kp1, d1 = ctxK.extract(img)
        s = synth.texture((240, 400), 6)
        ks, ds = small.extract(s)
        ref, rdesc, _ = O.orb(s, 300)
        assert_kp_equal(ks, ref); assert (ds == rdesc).all()
    finally:
        small.close()
    kp2, d2 = ctxK.extract(img)
    for kp, d in ((kp1, d1), (kp2, d2)):
        assert_kp_equal(kp, kp0); assert (d == d0).all()


# ---- opt-in projection windows in the batch path ----------------------------------
def test_batch_pass2_projection_windows_vs_oracle(ctxK):
    """map_win_uvr restricts every pass-2 row to the current keypoints under its window; the device gathers the
    candidates from a keypoint grid, the oracle scans every column with the same predicate: same claims."""
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    seq = synth.Sequence(seed=7)
    frames = [seq.frame(t) for t in range(3)]
    ext = [ctxK.extract(frames[t][0], cam=0) for t in range(3)]
    rng = np.random.default_rng(21)
    H, W = synth.K_SHAPE
    jobs = []
    for t in (1, 2):
        kp_prev, prev = ext[t - 1]
        kp_cur, cur = ext[t]
        M = 4000
        # map rows: descriptors of current keypoints with a few bits flipped (so windows decide who may match),
        # windows centred near the keypoint they came from, radius by octave; plus far-away, NaN, negative-radius,
        # image-covering and out-of-image windows
        src = rng.integers(0, len(cur), M)
        flips = np.packbits(rng.random((M, 256)) < 0.03, axis=1, bitorder="little")
        mp = cur[src] ^ flips
        uvr = np.stack([kp_cur["x"][src] + rng.normal(0, 6, M), kp_cur["y"][src] + rng.normal(0, 6, M),
                        15.0 * np.float32(1.2) ** kp_cur["octave"][src]], 1).astype(np.float32)
        far = rng.permutation(M)[:300]
        uvr[far, 0] = rng.uniform(0, W, 300); uvr[far, 1] = rng.uniform(0, H, 300)
        uvr[rng.permutation(M)[:20], 0] = np.nan
        uvr[rng.permutation(M)[:20], 2] = -1.0
        uvr[rng.permutation(M)[:5], 2] = 5000.0
        uvr[rng.permutation(M)[:20], 0] = -300.0
        uvr[rng.permutation(M)[:20], 1] = 4000.0
        mpr = np.full(M, -1, np.int32)
        take = rng.permutation(M)[:500]; psrc = rng.integers(0, len(prev), 500)
        mp[take] = prev[psrc]; mpr[take] = psrc
        live = (rng.random(len(prev)) < 0.8).astype(np.uint8)
        jobs.append(dict(left=frames[t][0], right=frames[t][1], bf=bf, baseline=b, prev_desc=prev, prev_live=live,
                         map_desc=mp, map_prev_row=mpr, map_win_uvr=uvr))
    ctxK.batch_submit(0, jobs); ctxK.batch_wait(0)
    total = 0
    for i, (job, t) in enumerate(zip(jobs, (1, 2))):
        r = ctxK.batch_result(0, i)
        kp_cur, cur = ext[t]
        assert (r["desc_left"] == cur).all()
        cxy = np.stack([kp_cur["x"], kp_cur["y"]], 1).astype(np.float32)
        p1 = O.match_greedy(job["prev_desc"], cur, 0, row_live=job["prev_live"])
        assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
        live2 = np.ones(len(job["map_desc"]), np.uint8)
        m = job["map_prev_row"] >= 0
        live2[m] = 1 - p1["row_claimed"][job["map_prev_row"][m]]
        p2 = O.match_greedy(job["map_desc"], cur, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                            row_base=len(job["prev_desc"]), win_uvr=job["map_win_uvr"], cur_xy=cxy)
        assert (r["p2_row_claimed"] == p2["row_claimed"]).all()
        assert (r["claim_row"] == p2["claim_row"]).all()
        total += int(p2["row_claimed"].sum())
        # and the windows matter: the brute-force scan claims a different set
        p2b = O.match_greedy(job["map_desc"], cur, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                             row_base=len(job["prev_desc"]))
        assert (p2b["claim_row"] != p2["claim_row"]).any()
    assert total > 300
    # the same frames without windows still take the brute-force path (graph keyed on the mode)
    plain = [{k: v for k, v in j.items() if k != "map_win_uvr"} for j in jobs]
    ctxK.batch_submit(0, plain); ctxK.batch_wait(0)
    r = ctxK.batch_result(0, 0)
    kp_cur, cur = ext[1]
    p1 = O.match_greedy(jobs[0]["prev_desc"], cur, 0, row_live=jobs[0]["prev_live"])
    live2 = np.ones(len(jobs[0]["map_desc"]), np.uint8)
    m = jobs[0]["map_prev_row"] >= 0
    live2[m] = 1 - p1["row_claimed"][jobs[0]["map_prev_row"][m]]
    p2b = O.match_greedy(jobs[0]["map_desc"], cur, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                         row_base=len(jobs[0]["prev_desc"]))
    assert (r["claim_row"] == p2b["claim_row"]).all()


def test_batch_windows_all_or_none(svo, ctxK):
    seq = synth.Sequence(seed=7)
    L, R = seq.frame(0)
    mp = np.zeros((10, 32), np.uint8)
    a = dict(left=L, right=R, bf=100.0, baseline=0.5, map_desc=mp, map_win_uvr=np.zeros((10, 3), np.float32))
    c = dict(left=L, right=R, bf=100.0, baseline=0.5, map_desc=mp)
    with pytest.raises(svo.SvoError) as e:
        ctxK.batch_submit(0, [a, c])
    assert e.value.code == svo.E_INVALID


# ---- opt-in quadtree ("octree") keypoint distribution (non-parity mode) -----------------
def octree_images(shape, seed):
    """Dense texture, texture confined to one corner (deep, unbalanced tree) and a featureless image."""
    h, w = shape
    dense = synth.texture(shape, seed)
    weak = 128 + (dense.astype(np.float32) - 128) * 0.3
    weak[: h // 3, : w // 4] = dense[: h // 3, : w // 4]
    return [dense, np.clip(np.rint(weak), 0, 255).astype(np.uint8), np.full(shape, 90, np.uint8)]


@pytest.mark.parametrize("shape,nf", [((240, 400), 500), ((376, 1241), 2000), ((203, 317), 300)])
def test_octree_extract_vs_oracle(svo, shape, nf):
    """svo_config.distribution = SVO_DIST_OCTREE: k_octree (sorted path codes, runs as nodes) against the oracle's
    explicit-node quadtree — same keypoints in the same order, same angles, responses and descriptors."""
    c = svo.Context(shape[1], shape[0], nfeatures=nf, max_batch=2, lanes=1, max_rows=1000, distribution=svo.DIST_OCTREE)
    try:
        imgs = octree_images(shape, 40 + nf)
        for img in imgs:
            ref, rdesc, _ = O.orb(img, nf, distribution=1)
            kp, desc = c.extract(img)
            assert_kp_equal(kp, ref)
            assert (desc == rdesc).all()
        assert len(c.extract(imgs[0])[0]) > nf // 2
        # the batch path runs the same stage
        ref0, d0, _ = O.orb(imgs[0], nf, distribution=1)
        ref1, d1, _ = O.orb(imgs[1], nf, distribution=1)
        c.batch_submit(0, [dict(left=imgs[0], right=imgs[1], bf=100.0, baseline=0.5),
                           dict(left=imgs[1], right=imgs[2], bf=100.0, baseline=0.5)])
        c.batch_wait(0)
        r = c.batch_result(0, 0)
        assert r["status"] == 0
        assert_kp_equal(r["kp_left"], ref0); assert_kp_equal(r["kp_right"], ref1)
        assert (r["desc_left"] == d0).all() and (r["desc_right"] == d1).all()
        r = c.batch_result(0, 1)
        assert_kp_equal(r["kp_left"], ref1); assert r["n_right"] == 0
    finally:
        c.close()


def test_octree_highres_uses_global_scratch(svo):
    """2560x720 / 8000: level 0 holds more corners than the shared-memory carve-out (12288 points), so the sort and
    the per-point state run from the level's global scratch arrays."""
    shape = (720, 2560)
    img = synth.texture(shape, 77)
    c = svo.Context(shape[1], shape[0], nfeatures=8000, max_batch=1, lanes=1, max_rows=1000, distribution=svo.DIST_OCTREE)
    try:
        kp, desc = c.extract(img)
        n0 = len(c.tap_list(0, svo.TAP_FAST, 0))
        assert n0 > 12288, n0
        ref, rdesc, _ = O.orb(img, 8000, distribution=1)
        assert_kp_equal(kp, ref)
        assert (desc == rdesc).all()
    finally:
        c.close()


def test_octree_mode_limits_and_default(svo, ctxK):
    with pytest.raises(svo.SvoError) as e:
        svo.Context(1241, 376, nfeatures=8000, nlevels=2, distribution=svo.DIST_OCTREE)   # level-0 quota 4364 > 4093 nodes
    assert e.value.code == svo.E_CAPACITY
    with pytest.raises(svo.SvoError) as e:
        svo.Context(1241, 376, nfeatures=2000, distribution=7)
    assert e.value.code == svo.E_INVALID
    # the default stays the parity path: an octree context changes nothing for a retainBest one
    img = synth.texture(synth.K_SHAPE, 9)
    ref, rdesc, _ = O.orb(img, 2000)
    kp, desc = ctxK.extract(img)
    assert_kp_equal(kp, ref); assert (desc == rdesc).all()


def test_batch_pass2_scans_only_the_columns_pass1_left_free(ctxK):
    """k_free_cols / the indexed tile of k_shortlist at their extremes, in one batch: every column claimed by pass 1
    (the previous frame IS the current one: nothing is left for pass 2), no column claimed (unrelated previous
    frame), and a frame without a previous frame; the oracle scans all columns and skips the claimed ones."""
    seq = synth.Sequence(seed=4)
    L, R = seq.frame(0)
    kl, dl, _ = O.orb(L, 2000)
    rng = np.random.default_rng(5)
    mp = synth.local_map([dl], rows=3000, seed=1)
    unrelated = rng.integers(0, 256, (1500, 32), dtype=np.uint8)
    near = dl[rng.permutation(len(dl))[:1200]].copy()
    near[:, 0] ^= 1                                            # one bit away: pass 1 claims 1200 distinct columns
    jobs = [dict(left=L, right=R, bf=100.0, baseline=0.5, prev_desc=dl.copy(), map_desc=mp),
            dict(left=L, right=R, bf=100.0, baseline=0.5, prev_desc=unrelated, map_desc=mp),
            dict(left=L, right=R, bf=100.0, baseline=0.5, prev_desc=near, map_desc=mp),
            dict(left=L, right=R, bf=100.0, baseline=0.5, map_desc=mp)]
    ctxK.batch_submit(0, jobs)
    ctxK.batch_wait(0)
    claimed_by_p1 = []
    for i, job in enumerate(jobs):
        r = ctxK.batch_result(0, i)
        assert r["status"] == 0 and (r["desc_left"] == dl).all()
        if "prev_desc" in job:
            p1 = O.match_greedy(job["prev_desc"], dl, 0)
            assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
            claimed, claim_row, base = p1["claimed"], p1["claim_row"], len(job["prev_desc"])
        else:
            claimed, claim_row, base = None, None, 0
        claimed_by_p1.append(0 if claimed is None else int(claimed.sum()))
        p2 = O.match_greedy(mp, dl, 1, claimed=claimed, claim_row=claim_row, row_base=base)
        assert (r["p2_row_claimed"] == p2["row_claimed"]).all(), i
        assert (r["claim_row"] == p2["claim_row"]).all(), i
    assert claimed_by_p1[0] == len(dl) and claimed_by_p1[1] == 0 and claimed_by_p1[2] == 1200 and claimed_by_p1[3] == 0
