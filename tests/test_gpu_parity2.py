"""GPU parity, part 2 (the holes round 1 left open): the pass-1 "dynamic" veto (src/pnpmatch.cc:101-144) through the
single-call and the batch API, the batch pipeline at BASELINE configs[2] (1241x376 / 4000 features) and configs[3]
(2560x720 / 8000 features) — where the free-column list no longer fits k_shortlist's tile, k_pairs runs more column
tiles and the resolver stages 9088 columns —, KITTI's other image shapes, and the limits the API states."""
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

# an F whose epipolar line for (lx, ly) is almost the horizontal line y = ly (stereo-like geometry), with awkward
# magnitudes so that the f64 arithmetic of the distance is not trivially exact
F_TEST = np.array([[1.1e-9, 2.3e-7, -3.1e-4], [-2.2e-7, 0.9e-9, 0.8312], [2.9e-4, -0.8297, 1.0]], np.float64)


@pytest.fixture(scope="module")
def svo():
    import svo as S
    return S


@pytest.fixture(scope="module")
def ctxK(svo):
    c = svo.Context(1241, 376, nfeatures=2000, max_batch=4, lanes=2, max_rows=5000)
    yield c
    c.close()


def noisy_copies(rng, base, n, ps=(0.0, 0.01, 0.03, 0.08, 0.2, 0.5)):
    src_idx = rng.integers(0, len(base), n)
    p = np.asarray(ps)[rng.integers(0, len(ps), n)]
    flips = np.packbits(rng.random((n, 256)) < p[:, None], axis=1, bitorder="little")
    return base[src_idx] ^ flips, src_idx


def line_y(F, lx, ly, cx):
    """y of the point with abscissa cx on the line F * (lx, ly, 1) (what src/pnpmatch.cc:110-113 measures against)."""
    A = F[0, 0] * lx + F[0, 1] * ly + F[0, 2]; B = F[1, 0] * lx + F[1, 1] * ly + F[1, 2]; C = F[2, 0] * lx + F[2, 1] * ly + F[2, 2]
    return -(A * cx + C) / B


def place_rows_near_epipolar_lines(rng, F, would_match, cur_xy, M, W=1241, H=376):
    """Row positions such that row i's would-be match lies 0.02 / 0.0999 / 0.1001 / 0.3 / 5 px off its epipolar line."""
    row_xy = np.stack([rng.uniform(0, W, M), rng.uniform(0, H, M)], 1)
    offs = np.array([0.02, 0.0999, 0.1001, 0.3, 5.0])
    for i in range(M):
        j = would_match[i]
        if j < 0:
            continue
        cx, cy = float(cur_xy[j, 0]), float(cur_xy[j, 1])
        lx = rng.uniform(0, W)
        # solve for ly so that the line passes `off` px from (cx, cy): a few fixed-point steps are plenty
        off = offs[rng.integers(0, len(offs))] * (1 if rng.random() < 0.5 else -1)
        ly = cy
        for _ in range(6):
            ly += (cy + off) - line_y(F, lx, ly, cx)
        row_xy[i] = (lx, ly)
    return row_xy.astype(np.float32)


def test_match_greedy_veto_vs_oracle(ctxK):
    """svo_match_greedy with svo_veto: rows whose would-be match falls inside a box and off the epipolar line mark
    the map point bad and claim nothing, which changes what later rows see."""
    rng = np.random.default_rng(31)
    N, M = 1800, 2500
    cur = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    rows, _ = noisy_copies(rng, cur, M)
    cur_xy = np.stack([rng.uniform(0, 1241, N), rng.uniform(0, 376, N)], 1).astype(np.float32)
    live = (rng.random(M) < 0.9).astype(np.uint8)
    plain = O.match_greedy(rows, cur, 0, row_live=live)
    row_xy = place_rows_near_epipolar_lines(rng, F_TEST, plain["best_idx"], cur_xy, M)
    boxes = np.array([[100, 500, 50, 200], [700, 1200, 100, 376], [0, 60, 0, 376]], np.int32)
    veto = dict(boxes=boxes, F=F_TEST, row_xy=row_xy, cur_xy=cur_xy)
    ref = O.match_greedy(rows, cur, 0, row_live=live, veto=veto, row_base=5)
    got = ctxK.match_greedy(rows, cur, 0, row_live=live, veto=veto, row_base=5)
    for k in ("row_bad", "row_claimed", "claimed", "claim_row"):
        assert (ref[k] == got[k]).all(), k
    lv = live.astype(bool)
    for k in ("best_idx", "best", "second"):
        assert (ref[k][lv] == got[k][lv]).all(), k
    nb = int(ref["row_bad"].sum())
    assert 50 < nb < plain["row_claimed"].sum(), nb
    assert (ref["claim_row"] != plain["claim_row"] + np.where(plain["claim_row"] >= 0, 5, 0)).any()
    # no boxes / veto without F: identical to the plain scan
    none = ctxK.match_greedy(rows, cur, 0, row_live=live, veto=dict(boxes=np.zeros((0, 4), np.int32), F=F_TEST, row_xy=row_xy, cur_xy=cur_xy))
    assert (none["row_claimed"] == plain["row_claimed"]).all() and not none["row_bad"].any()
    # the veto belongs to pass 1 only
    p2 = ctxK.match_greedy(rows, cur, 1, row_live=live, veto=veto)
    r2 = O.match_greedy(rows, cur, 1, row_live=live)
    assert (p2["row_claimed"] == r2["row_claimed"]).all() and not p2["row_bad"].any()


def batch_oracle(job, dl, kp_left):
    """The matchers of one batch frame on the oracle: BF, pass 1 (with veto), pass 2 (skipping linked rows that
    pass 1 matched or marked bad)."""
    out = {}
    prev = job.get("prev_desc")
    if prev is not None:
        out["bf"] = O.match_bf(dl, prev)
        veto = None
        if job.get("boxes") is not None:
            cxy = np.stack([kp_left["x"], kp_left["y"]], 1).astype(np.float32)
            veto = dict(boxes=job["boxes"], F=job["F"], row_xy=job["prev_xy"], cur_xy=cxy)
        p1 = O.match_greedy(prev, dl, 0, row_live=job.get("prev_live"), veto=veto)
        out["p1"] = p1
        claimed, claim_row, base = p1["claimed"], p1["claim_row"], len(prev)
    else:
        claimed, claim_row, base = None, None, 0
    mp = job.get("map_desc")
    if mp is not None:
        live2 = np.ones(len(mp), np.uint8)
        mpr = job.get("map_prev_row")
        if mpr is not None and prev is not None:
            m = (mpr >= 0) & (mpr < len(prev))
            live2[m] = 1 - (p1["row_claimed"][mpr[m]] | p1["row_bad"][mpr[m]])
        out["p2"] = O.match_greedy(mp, dl, 1, claimed=claimed, claim_row=claim_row, row_live=live2, row_base=base)
    return out


def check_batch_frame(r, job, ora):
    kl, dl, kr, dr, ur, dep, mr, sad = ora
    assert r["status"] == 0
    assert len(r["kp_left"]) == len(kl) and len(r["kp_right"]) == len(kr)
    for f in ("x", "y", "angle", "response"):
        assert (r["kp_left"][f].view(np.uint32) == kl[f].view(np.uint32)).all(), f
        assert (r["kp_right"][f].view(np.uint32) == kr[f].view(np.uint32)).all(), f
    assert (r["kp_left"]["octave"] == kl["octave"]).all()
    assert (r["desc_left"] == dl).all() and (r["desc_right"] == dr).all()
    valid = dep > 0
    assert ((r["depth"] > 0) == valid).all() and r["n_stereo"] == valid.sum()
    if valid.any():
        assert np.abs(r["u_right"][valid] - ur[valid]).max() <= 1e-3
        assert (np.abs(r["depth"][valid] - dep[valid]) <= 1e-3 * np.abs(dep[valid])).all()
    m = batch_oracle(job, dl, kl)
    if "bf" in m:
        oi, od, ok = m["bf"]
        assert (r["bf_idx"] == oi).all() and (r["bf_dist"] == od).all() and (r["bf_keep"] == ok).all()
        p1 = m["p1"]
        assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
        if job.get("boxes") is not None:
            assert (r["p1_row_bad"] == p1["row_bad"]).all()
        lv = job["prev_live"].astype(bool) if job.get("prev_live") is not None else np.ones(len(p1["best"]), bool)
        for k in ("best_idx", "best", "second"):
            assert (r["p1_" + k][lv] == p1[k][lv]).all(), k
    if "p2" in m:
        assert (r["p2_row_claimed"] == m["p2"]["row_claimed"]).all()
        assert (r["claim_row"] == m["p2"]["claim_row"]).all()
    return m


def stereo_oracle(L, R, nf, bf, b):
    kl, dl, pl = O.orb(L, nf, with_pyramid=True)
    kr, dr, pr = O.orb(R, nf, with_pyramid=True)
    ur, dep, mr, sad = O.stereo_sparse(kl, dl, pl, kr, dr, pr, bf, b)
    O.pyramid_free(pl); O.pyramid_free(pr)
    return kl, dl, kr, dr, ur, dep, mr, sad


def make_jobs(frames, ora, rows, rng, n_link, with_veto):
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    jobs = []
    for t in range(1, len(frames)):
        prev_desc, kp_prev = ora[t - 1][1], ora[t - 1][0]
        live = (ora[t - 1][5] > 0).astype(np.uint8)
        live |= (rng.random(len(live)) < 0.5).astype(np.uint8)
        mp = synth.local_map([o[1] for o in ora[:t]], rows=rows, seed=t)
        mpr = np.full(rows, -1, np.int32)
        take = rng.permutation(rows)[:n_link]
        src = rng.integers(0, len(prev_desc), n_link)
        mp[take] = prev_desc[src]; mpr[take] = src
        mpr[rng.permutation(rows)[:7]] = len(prev_desc) + 3      # links past the previous set are no links (ADVICE)
        job = dict(left=frames[t][0], right=frames[t][1], bf=bf, baseline=b, prev_desc=prev_desc, prev_live=live,
                   map_desc=mp, map_prev_row=mpr)
        if with_veto:
            kl, dl = ora[t][0], ora[t][1]
            cxy = np.stack([kl["x"], kl["y"]], 1).astype(np.float32)
            plain = O.match_greedy(prev_desc, dl, 0, row_live=live)
            job["prev_xy"] = place_rows_near_epipolar_lines(rng, F_TEST, plain["best_idx"], cxy, len(prev_desc),
                                                            W=frames[t][0].shape[1], H=frames[t][0].shape[0])
            W, H = frames[t][0].shape[1], frames[t][0].shape[0]
            job["boxes"] = np.array([[W // 10, W // 2, H // 8, H // 2], [W // 2 + 40, W - 30, H // 4, H]], np.int32)
            job["F"] = F_TEST.copy()
        jobs.append(job)
    return jobs


def test_batch_veto_vs_oracle(ctxK):
    """svo_frame_in.boxes / F / prev_xy: the fused pass-1 resolver applies the veto, p1_row_bad comes back, and pass 2
    skips the local-map rows linked to a row that turned bad (mp_local->bad, src/pnpmatch.cc:163)."""
    cal = synth.KITTI_04_12
    bf, b = np.float32(cal["bf"]), np.float32(cal["bf"] / cal["fx"])
    seq = synth.Sequence(seed=12)
    frames = [seq.frame(t) for t in range(3)]
    ora = [stereo_oracle(L, R, 2000, bf, b) for L, R in frames]
    rng = np.random.default_rng(19)
    jobs = make_jobs(frames, ora, 5000, rng, 1500, True)
    plain_jobs = [{k: v for k, v in j.items() if k not in ("boxes", "F", "prev_xy")} for j in jobs]
    ctxK.batch_submit(0, jobs); ctxK.batch_submit(1, plain_jobs)
    ctxK.batch_wait(0); ctxK.batch_wait(1)
    for i, job in enumerate(jobs):
        r = ctxK.batch_result(0, i)
        m = check_batch_frame(r, job, ora[i + 1])
        nbad = int(m["p1"]["row_bad"].sum())
        assert nbad > 20, nbad
        # at least one bad row is linked from the map, so the skip in pass 2 is exercised
        mpr = job["map_prev_row"]
        ok = (mpr >= 0) & (mpr < len(job["prev_desc"]))
        assert m["p1"]["row_bad"][mpr[ok]].sum() > 0
        # the same frame without veto inputs gives the un-vetoed result (other graph, same lane buffers)
        rp = ctxK.batch_result(1, i)
        check_batch_frame(rp, plain_jobs[i], ora[i + 1])
        assert (rp["claim_row"] != r["claim_row"]).any()
    # mixed batch: one frame with veto inputs, one without, one frame without a previous frame
    mixed = [jobs[0], plain_jobs[1], dict(left=frames[1][0], right=frames[1][1], bf=float(bf), baseline=float(b),
                                          map_desc=jobs[0]["map_desc"])]
    ctxK.batch_submit(0, mixed); ctxK.batch_wait(0)
    check_batch_frame(ctxK.batch_result(0, 0), mixed[0], ora[1])
    r1 = ctxK.batch_result(0, 1)
    check_batch_frame(r1, mixed[1], ora[2])
    assert not r1["p1_row_bad"].any()
    check_batch_frame(ctxK.batch_result(0, 2), mixed[2], ora[1])


def test_batch_veto_rejects_bad_arguments(svo, ctxK):
    L, R = synth.Sequence(seed=7).frame(0)
    base = dict(left=L, right=R, bf=100.0, baseline=0.5, prev_desc=np.zeros((10, 32), np.uint8),
                prev_xy=np.zeros((10, 2), np.float32), F=F_TEST)
    with pytest.raises(svo.SvoError) as e:
        ctxK.batch_submit(0, [dict(base, boxes=np.zeros((257, 4), np.int32))])
    assert e.value.code == svo.E_INVALID
    with pytest.raises(svo.SvoError) as e:
        ctxK.match_greedy(np.zeros((4, 32), np.uint8), np.zeros((4, 32), np.uint8), 0,
                          veto=dict(boxes=np.zeros((257, 4), np.int32), F=F_TEST, row_xy=np.zeros((4, 2)), cur_xy=np.zeros((4, 2))))
    assert e.value.code == svo.E_CAPACITY


@pytest.mark.parametrize("shape,nf,rows,veto", [(synth.K_SHAPE, 4000, 5000, False), (synth.H_SHAPE, 8000, 7000, True)],
                         ids=["K4000", "H8000"])
def test_batch_pipeline_other_baseline_configs(svo, shape, nf, rows, veto):
    """The whole batch path (extraction, stereo, BF, pass 1, pass 2) at BASELINE configs[2] and configs[3]."""
    cal = synth.KITTI_04_12
    bf, b = np.float32(cal["bf"]), np.float32(cal["bf"] / cal["fx"])
    H, W = shape
    seq = synth.Sequence(shape, seed=3)
    frames = [seq.frame(t) for t in range(3)]
    ora = [stereo_oracle(L, R, nf, bf, b) for L, R in frames]
    rng = np.random.default_rng(23)
    jobs = make_jobs(frames, ora, rows, rng, rows // 5, veto)
    c = svo.Context(W, H, nfeatures=nf, max_batch=2, lanes=2, max_rows=max(rows, nf + nf // 8 + 128))   # rows of a set: map or previous frame
    try:
        c.batch_submit(0, jobs)
        c.batch_submit(1, jobs[1:])                     # a one-frame batch on the other lane at the same time
        c.batch_wait(0); c.batch_wait(1)
        tot1 = tot2 = 0
        for i, job in enumerate(jobs):
            m = check_batch_frame(c.batch_result(0, i), job, ora[i + 1])
            tot1 += int(m["p1"]["row_claimed"].sum()); tot2 += int(m["p2"]["row_claimed"].sum())
            free = int((m["p1"]["claimed"] == 0).sum())
            assert free > 3296 or nf < 8000, free          # H/8000: the free columns exceed k_shortlist's tile
        check_batch_frame(c.batch_result(1, 0), jobs[1], ora[2])
        assert tot1 > 200 and tot2 > 100, (tot1, tot2)
        # graph replay of the same (n, stages) key with other inputs
        c.batch_submit(0, jobs[::-1]); c.batch_wait(0)
        check_batch_frame(c.batch_result(0, 0), jobs[1], ora[2])
        check_batch_frame(c.batch_result(0, 1), jobs[0], ora[1])
    finally:
        c.close()


@pytest.mark.parametrize("shape", [(375, 1242), (370, 1226)], ids=["1242x375", "1226x370"])
def test_other_kitti_shapes(svo, shape):
    """KITTI's other rectified sizes (sequences 00-02 are 1241x376, 03 is 1242x375, 04-12 are 1226x370)."""
    H, W = shape
    L, R, _ = synth.stereo_pair(shape, seed=13)
    c = svo.Context(W, H, nfeatures=2000, max_batch=1, lanes=1, max_rows=2500)
    try:
        for cam, img in enumerate((L, R)):
            kp, desc = c.extract(img, cam=cam)
            ref, rdesc, _ = O.orb(img, 2000)
            assert len(kp) == len(ref) and (desc == rdesc).all()
            for f in ("x", "y", "angle", "response"):
                assert (kp[f].view(np.uint32) == ref[f].view(np.uint32)).all(), f
        cal = synth.KITTI_04_12
        bf, b = np.float32(cal["bf"]), np.float32(cal["bf"] / cal["fx"])
        kl, dl, kr, dr, ur, dep, mr, sad = stereo_oracle(L, R, 2000, bf, b)
        gur, gdep, gmr, gsad = c.stereo_sparse(float(bf), float(b))
        valid = dep > 0
        assert valid.sum() > 200 and ((gdep > 0) == valid).all()
        assert (gmr[valid] == mr[valid]).all() and np.abs(gur[valid] - ur[valid]).max() <= 1e-3
    finally:
        c.close()


def test_stated_limits(svo):
    """The limits include/svo_b200.h and svo_create state are enforced with a status, not a crash."""
    for kw, code in ((dict(width=4096, height=400), svo.E_INVALID),          # images up to 4095 x 4095
                     (dict(width=400, height=4096), svo.E_INVALID),
                     (dict(nfeatures=30000), svo.E_INVALID),                 # matcher limit on keypoints per frame
                     (dict(max_rows=40001), svo.E_INVALID),                  # rows per greedy set
                     (dict(width=63, height=63), svo.E_INVALID),
                     (dict(nlevels=9), svo.E_INVALID)):
        args = dict(width=640, height=480, nfeatures=500, max_rows=1000)
        args.update(kw)
        with pytest.raises(svo.SvoError) as e:
            svo.Context(**args)
        assert e.value.code == code, (kw, e.value.code)
    # the largest supported keypoint count still creates and extracts (4095-wide image, 16000 features)
    c = svo.Context(4095, 600, nfeatures=16000, max_batch=1, lanes=1, max_rows=1000)
    try:
        img = synth.texture((600, 4095), 3)
        kp, desc = c.extract(img, cap=40000)
        ref, rdesc, _ = O.orb(img, 16000)
        assert len(kp) == len(ref) and (desc == rdesc).all()
        # capacity errors of the single-call matchers
        with pytest.raises(svo.SvoError) as e:
            c.match_bf(np.zeros((20000, 32), np.uint8), np.zeros((10, 32), np.uint8))
        assert e.value.code == svo.E_CAPACITY
    finally:
        c.close()


# ---- opt-in projection windows computed on the device ("true projection-guided windows") ----
def predicted_pose():
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = synth.rodrigues([0.01, -0.03, 0.005]).astype(np.float32); T[:3, 3] = [0.2, -0.05, 0.6]
    return T


def map_points_for(kp, depth_guess, Tcw, K4, rng):
    """World points that project (under Tcw) near the given keypoints: back-project at a guessed depth, undo the pose."""
    fx, fy, cx, cy = K4
    z = depth_guess
    Xc = np.stack([(kp["x"] - cx) / fx * z, (kp["y"] - cy) / fy * z, z], 1).astype(np.float64)
    R, t = Tcw[:3, :3].astype(np.float64), Tcw[:3, 3].astype(np.float64)
    Xw = (Xc - t) @ R                   # R^T (Xc - t)
    return (Xw + rng.normal(0, 0.02, Xw.shape)).astype(np.float32)


def test_project_map_vs_oracle(ctxK):
    rng = np.random.default_rng(41)
    cal = synth.KITTI_04_12
    K4 = tuple(np.float32(cal[k]) for k in ("fx", "fy", "cx", "cy"))
    T = predicted_pose()
    n = 6000
    xyz = np.stack([rng.uniform(-30, 30, n), rng.uniform(-8, 8, n), rng.uniform(-5, 80, n)], 1).astype(np.float32)
    xyz[:50, 2] = 0.0; xyz[50:60] = 0.0                      # on / behind the camera plane
    octv = rng.integers(-1, 10, n).astype(np.int32)            # out-of-range levels are clamped
    ref = O.project_map(xyz, octv, T, K4, 1241, 376, th=7.0)
    got = ctxK.project_map(xyz, octv, T, K4, th=7.0)
    assert (got.view(np.uint32) == ref.view(np.uint32)).all()
    vis = ref[:, 2] > 0
    assert 500 < vis.sum() < n - 500
    assert (ref[~vis] == np.array([0, 0, -1], np.float32)).all()
    # independent numpy float32 restatement of the formula for the visible points
    X = xyz[vis]
    xc = ((T[0, 0] * X[:, 0] + T[0, 1] * X[:, 1]) + T[0, 2] * X[:, 2]) + T[0, 3]
    zc = ((T[2, 0] * X[:, 0] + T[2, 1] * X[:, 1]) + T[2, 2] * X[:, 2]) + T[2, 3]
    u = (K4[0] * xc) * (np.float32(1) / zc) + K4[2]
    assert (u.astype(np.float32).view(np.uint32) == ref[vis, 0].view(np.uint32)).all()
    assert (ctxK.project_map(xyz, None, T, K4)[vis, 2] == np.float32(7.0)).all()


def test_batch_pass2_projected_windows_vs_oracle(ctxK):
    """svo_frame_in.map_xyz / map_octave / Tcw_pred: the windows are projected on the device inside the batch and pass 2
    gathers from them; same claims as the oracle's windowed scan over the oracle's own projection."""
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    K4 = tuple(np.float32(cal[k]) for k in ("fx", "fy", "cx", "cy"))
    seq = synth.Sequence(seed=14)
    frames = [seq.frame(t) for t in range(3)]
    ext = [ctxK.extract(frames[t][0], cam=0) for t in range(3)]
    rng = np.random.default_rng(33)
    T = predicted_pose()
    jobs = []
    for t in (1, 2):
        kp_prev, prev = ext[t - 1]
        kp_cur, cur = ext[t]
        M = 3500
        src = rng.integers(0, len(cur), M)
        mp = cur[src] ^ np.packbits(rng.random((M, 256)) < 0.03, axis=1, bitorder="little")
        xyz = map_points_for(kp_cur[src], rng.uniform(4, 60, M).astype(np.float32), T, K4, rng)
        far = rng.permutation(M)[:400]
        xyz[far] = np.stack([rng.uniform(-30, 30, 400), rng.uniform(-8, 8, 400), rng.uniform(-5, 80, 400)], 1).astype(np.float32)
        jobs.append(dict(left=frames[t][0], right=frames[t][1], bf=bf, baseline=b, prev_desc=prev, map_desc=mp,
                         map_xyz=xyz, map_octave=kp_cur["octave"][src].astype(np.int32), Tcw_pred=T, K=K4, proj_th=7.0))
    ctxK.batch_submit(0, jobs); ctxK.batch_wait(0)
    total = 0
    for i, (job, t) in enumerate(zip(jobs, (1, 2))):
        r = ctxK.batch_result(0, i)
        kp_cur, cur = ext[t]
        assert (r["desc_left"] == cur).all()
        cxy = np.stack([kp_cur["x"], kp_cur["y"]], 1).astype(np.float32)
        win = O.project_map(job["map_xyz"], job["map_octave"], T, K4, 1241, 376, th=7.0)
        p1 = O.match_greedy(job["prev_desc"], cur, 0)
        assert (r["p1_row_claimed"] == p1["row_claimed"]).all()
        p2 = O.match_greedy(job["map_desc"], cur, 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_base=len(job["prev_desc"]),
                            win_uvr=win, cur_xy=cxy)
        assert (r["p2_row_claimed"] == p2["row_claimed"]).all()
        assert (r["claim_row"] == p2["claim_row"]).all()
        total += int(p2["row_claimed"].sum())
    assert total > 100, total


def test_skip_match_score_changes_nothing_else(svo):
    """svo_config.skip_match_score: the (bestIdx, best, second) triplet of src/pnpmatch.cc:99 is not produced; claims,
    BF matches and everything else are byte-identical to the default context's."""
    cal = synth.KITTI_04_12
    bf, b = float(np.float32(cal["bf"])), float(np.float32(cal["bf"] / cal["fx"]))
    seq = synth.Sequence(seed=15)
    frames = [seq.frame(t) for t in range(2)]
    res = []
    for skip in (False, True):
        c = svo.Context(1241, 376, nfeatures=2000, max_batch=2, lanes=1, max_rows=5000, skip_match_score=skip)
        try:
            prev = c.extract(frames[0][0])[1]
            mp = synth.local_map([prev], rows=3000, seed=2)
            c.batch_submit(0, [dict(left=frames[1][0], right=frames[1][1], bf=bf, baseline=b, prev_desc=prev, map_desc=mp)])
            c.batch_wait(0)
            res.append(c.batch_result(0, 0))
        finally:
            c.close()
    full, lean = res
    assert full["p1_best"].min() < 15 and (lean["p1_best"] == 0).all()          # absent: the binding shows zeros
    for k in full:
        if k.startswith("p1_best") or k == "p1_second":
            continue
        a, bb = full[k], lean[k]
        assert (a.tobytes() == bb.tobytes()) if isinstance(a, np.ndarray) else a == bb, k


@pytest.mark.parametrize("mode", ["1", "0"])
def test_both_pyramid_forms_are_byte_exact(svo, mode):
    """SVO_B200_PYRAMID_FUSED=1: the one-launch pyramid (TMA bulk copy of the band's level-0 rows, levels 1..7 chained in
    shared memory; the default for launches of up to 8 images) and =0: the seven per-level launches (the default above
    that) give the same bytes as the oracle on every level, for a KITTI-shape, an odd-sized and a high-resolution image,
    and the whole extractor on top of them is unchanged."""
    import os
    os.environ["SVO_B200_PYRAMID_FUSED"] = mode
    try:
        for shape, nf in (((376, 1241), 2000), ((203, 317), 300), ((720, 2560), 4000)):
            img = synth.texture(shape, 7 + nf)
            c = svo.Context(shape[1], shape[0], nfeatures=nf, max_batch=1, lanes=1, max_rows=1000)
            try:
                kp, desc = c.extract(img)
                ref, rdesc, pyr = O.orb(img, nf, with_pyramid=True)
                for l in range(8):
                    assert (c.tap_image(0, l) == pyr.level(l)).all(), (shape, l)
                O.pyramid_free(pyr)
                assert len(kp) == len(ref) and (desc == rdesc).all()
            finally:
                c.close()
    finally:
        del os.environ["SVO_B200_PYRAMID_FUSED"]


@pytest.mark.parametrize("band", ["8", "16"])
def test_both_fast_band_heights_give_the_same_corners(svo, band):
    """k_fast works on bands of 8 rows (latency contexts, very wide images) or 16 rows (batch contexts: max_batch >= 8);
    SVO_B200_FAST_BAND forces one.  Either way the raster-ordered FAST list of every level, and the extractor on top
    of it, equal the oracle's — for a KITTI-shape, an odd-sized and a high-resolution image (whose band rows are not
    a multiple of either height) and through the batch path of a max_batch = 8 context."""
    import os
    os.environ["SVO_B200_FAST_BAND"] = band
    try:
        for shape, nf in (((376, 1241), 2000), ((203, 317), 300), ((720, 2560), 4000)):
            img = synth.texture(shape, 11 + nf)
            c = svo.Context(shape[1], shape[0], nfeatures=nf, max_batch=1, lanes=1, max_rows=1000)
            try:
                kp, desc = c.extract(img)
                ref, rdesc, pyr = O.orb(img, nf, with_pyramid=True)
                for l in range(8):
                    xs, ys, sc = O.fast_nms(pyr.level(l), 20, 31)
                    f = c.tap_list(0, svo.TAP_FAST, l)
                    assert len(f) == len(xs) and (f[:, 0] == xs).all() and (f[:, 1] == ys).all() and (f[:, 2] == sc).all(), (shape, l)
                O.pyramid_free(pyr)
                assert len(kp) == len(ref) and (desc == rdesc).all()
            finally:
                c.close()
    finally:
        del os.environ["SVO_B200_FAST_BAND"]
    # the default choice of a batch context (16 rows) through svo_batch_submit
    c = svo.Context(400, 240, nfeatures=500, max_batch=8, lanes=1, max_rows=1000)
    try:
        seq = synth.Sequence((240, 400), seed=3)
        frames = [seq.frame(t) for t in range(8)]
        c.batch_submit(0, [dict(left=f[0], right=f[1], bf=379.8145, baseline=0.5372) for f in frames]); c.batch_wait(0)
        for i, f in enumerate(frames):
            r = c.batch_result(0, i)
            kl, dl, _ = O.orb(f[0], 500)
            assert r["n_left"] == len(kl) and (r["desc_left"] == dl).all() and (r["kp_left"]["x"] == kl["x"]).all()
    finally:
        c.close()


@pytest.mark.parametrize("var", ["SVO_B200_RESIZE_QUADS", "SVO_B200_HARRIS8", "SVO_B200_BLUR_MARGIN", "SVO_B200_BLUR_PACK"])
def test_previous_kernel_forms_stay_exact(svo, var):
    """The kernels the instruction-count work replaced stay in the library behind switches (k_resize for levels whose taps do
    not fit k_resize_q's window, k_harris, k_blur<false>, evenly split blur tiles): with one switched back, pyramid and blur
    bytes, the Harris-ranked second cull and the extractor on top equal the oracle's, for a KITTI-shape and an odd-sized image."""
    import os
    os.environ[var] = "0"
    try:
        for shape, nf in (((376, 1241), 2000), ((203, 317), 300)):
            img = synth.texture(shape, 23 + nf)
            c = svo.Context(shape[1], shape[0], nfeatures=nf, max_batch=1, lanes=1, max_rows=1000)
            try:
                kp, desc = c.extract(img)
                ref, rdesc, pyr = O.orb(img, nf, with_pyramid=True)
                for l in range(8):
                    assert (c.tap_image(0, l) == pyr.level(l)).all(), (var, shape, l)
                    lh, lw = pyr.level(l).shape
                    if lh > 62 and lw > 62:      # the oracle blurs only levels that can hold a keypoint (31-px border on each side)
                        assert (c.tap_image(0, l, blurred=True) == pyr.level(l, True)).all(), (var, shape, l)
                O.pyramid_free(pyr)
                assert len(kp) == len(ref) and (desc == rdesc).all()
                for f in ("x", "y", "angle", "response", "octave"):
                    assert (kp[f] == ref[f]).all(), (var, shape, f)
            finally:
                c.close()
    finally:
        del os.environ[var]


@pytest.mark.parametrize("nl,sf,nf,shape", [(5, 1.3, 800, (376, 1241)), (3, 1.5, 300, (240, 400)), (8, 1.1, 1500, (376, 1241)),
                                            (1, 1.2, 200, (240, 400)), (6, 2.0, 400, (480, 640))])
def test_other_level_counts_and_scale_factors(svo, nl, sf, nf, shape):
    """ORBextractor.nLevels / scaleFactor other than KITTI's 8 / 1.2 (the yaml files are the reference's only source of
    them): the extractor still equals the oracle bit for bit, with 1 to 8 levels and scale factors from 1.1 to 2."""
    img = synth.texture(shape, 17 + nl)
    c = svo.Context(shape[1], shape[0], nfeatures=nf, nlevels=nl, scale_factor=sf, max_batch=1, lanes=1, max_rows=500)
    try:
        kp, desc = c.extract(img)
        okp, odesc, _ = O.orb(img, nf, scale=sf, nlevels=nl)
        assert len(kp) == len(okp) and (desc == odesc).all() and (kp["octave"] == okp["octave"]).all()
        for f in ("x", "y", "angle", "response", "size"):
            assert (kp[f].view(np.uint32) == okp[f].view(np.uint32)).all(), f
    finally:
        c.close()


@pytest.mark.parametrize("shape,nf", [((181, 249), 200), ((297, 465), 600)])
def test_extract_at_sizes_whose_level_rounding_sits_on_a_half(svo, shape, nf):
    """249 / 1.2 = 207.5 - 8e-6: cv2 4.13 makes level 1 208 wide (cvRound of a float multiplication by the reciprocal of
    the scale), the quotient rounds to 207 (tests/test_oracle_vs_cv2.py).  svo_extract against cv2's recorded output
    (tests/golden/half_sizes.npz, make_golden_half_sizes.py) and the oracle; level geometry against the oracle's."""
    h, w = shape
    img = synth.texture(shape, 77)
    c = svo.Context(w, h, nfeatures=nf, max_rows=1000)
    try:
        lw, lh, ls, q = c.geometry()
        olw, olh, ols, oq = O.geometry(w, h, 8, 1.2, nf)
        assert (lw == olw).all() and (lh == olh).all() and (q == oq).all()
        assert lw[1] == {249: 208, 465: 388}[w] and lh[1] == {181: 151, 297: 248}[h]
        kp, desc = c.extract(img)
    finally:
        c.close()
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "half_sizes.npz"))
    for rkp, rdesc in ((g["kp_%dx%d" % (w, h)], g["desc_%dx%d" % (w, h)]), O.orb(img, nf)[:2]):
        assert len(kp) == len(rkp) > 0 and (desc == rdesc).all() and (kp["octave"] == rkp["octave"]).all()
        for f in ("x", "y", "size", "angle", "response"):
            assert (kp[f].view(np.uint32) == rkp[f].view(np.uint32)).all(), f
