"""Pins oracle/track.py — the CPU restatement of what Tracking::Track carries from frame to frame (the oracle of the
device-resident tracker state, csrc/track.cu) — to the REFERENCE'S OWN CODE: oracle/_ref drives the reference's frame,
mappoint and pnpmatch classes for several frames as Tracking::Track does (oracle/ref.py:run_sequence) and every
frame's CurrentFrame->MapPoints, match_score, bad flags, created points and local map are compared.  Needs
/root/reference (or a prebuilt oracle/_ref); the committed fixture tests/golden/track_seq.npz carries the same
expectations to the GPU box."""
import os

import numpy as np
import pytest

import synth
from oracle import ref as R
from oracle import track as T

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
K4 = (CAL["fx"], CAL["fy"], CAL["cx"], CAL["cy"])
BF = np.float32(CAL["bf"])
BOXES = [[300, 700, 100, 300], [900, 1100, 50, 200], [20, 180, 200, 360]]

needs_ref = pytest.mark.skipif(not R.available(), reason="/root/reference (or a prebuilt oracle/_ref) is not present")


def reference_run(seed, n, boxes_of):
    seq = synth.Sequence(synth.K_SHAPE, seed=seed)
    frames = [seq.frame(t) for t in range(n)]
    disps = [synth.dense_disparity(synth.K_SHAPE, 100 * seed + t) for t in range(n)]
    return R.run_sequence(frames, disps, K, BF, [boxes_of(t) for t in range(n)])


def good_names(m):
    """(create_id, idx) of the set's points in its own order, without the bad ones: the reference leaves a point the
    veto marked bad in LocalMapPoints until it ages out but never uses it again (src/pnpmatch.cc:66,163); the tracker
    state drops it at once."""
    return [(c, i) for c, i, b in zip(m["create_id"].tolist(), m["idx"].tolist(), m["bad"].tolist()) if not b]


def replay(recs, boxes_of, impose_order=True):
    """oracle/track.py over the per-frame inputs the reference saw; returns per-frame outputs."""
    trk = T.Tracker(window=4)
    outs = []
    for t, rec in enumerate(recs):
        cur = rec["cur"]
        if impose_order:   # pass 2 walks the std::set in pointer order: impose the reference's own order
            trk.reorder(good_names(rec["map_before"]))
        o = trk.step(cur["kps"][:, :2], cur["desc"], cur["depth_at_kp"], t, boxes=boxes_of(t), F=rec["F"], K4=K4)
        outs.append((o, trk.names()))
    return outs


def check(recs, outs):
    n_bad = n_p1 = n_p2 = 0
    for t, (rec, (o, names)) in enumerate(zip(recs, outs)):
        cur, last = rec["cur"], rec["last"]
        N = cur["N"]
        assert len(cur["kps"]) == N == 500
        # the local map the frame saw
        assert o["n_map"] == len(good_names(rec["map_before"]))
        if t > 0:
            # CurrentFrame->MapPoints after both passes (before createmappoint): only matched keypoints own a point
            matched = o["claim_row"] >= 0
            got_c = np.where(matched, o["mp_create"], -1); got_i = np.where(matched, o["mp_idx"], -1)
            assert (cur["mp_create_id"] == got_c).all() and (cur["mp_idx"] == got_i).all(), "frame %d" % t
            # match_score (src/pnpmatch.cc:99) of the live pass-1 rows, -1 elsewhere
            lv = recs[t - 1]["last"]["mp_create_id"] >= 0
            with np.errstate(divide="ignore", invalid="ignore"):
                score = o["p1_second"].astype(np.float32) / o["p1_best"].astype(np.float32)
            assert (cur["match_score"][lv].view(np.uint32) == score[lv].view(np.uint32)).all()
            assert (cur["match_score"][~lv] == -1).all()
            # mp->bad as the veto left it on the last frame's points
            assert (rec["last_after_match"]["mp_bad"][lv] == o["p1_row_bad"][lv]).all()
            n_bad += int(o["p1_row_bad"].sum()); n_p1 += int(o["p1_row_claimed"].sum()); n_p2 += int(o["p2_row_claimed"].sum())
        # lastframe after createmappoint: every keypoint's point
        assert rec["created"] == o["created"]
        assert (last["mp_create_id"] == o["mp_create"]).all() and (last["mp_idx"] == o["mp_idx"]).all(), "frame %d" % t
        # the local map after createmappoint and the 4-frame window, as a set
        assert sorted(good_names(rec["map"])) == sorted(names), "frame %d" % t
    return n_bad, n_p1, n_p2


@needs_ref
def test_tracker_oracle_follows_the_reference_over_seven_frames():
    boxes_of = lambda t: BOXES if t % 2 == 1 else BOXES[:1]
    recs = reference_run(5, 7, boxes_of)
    outs = replay(recs, boxes_of)
    n_bad, n_p1, n_p2 = check(recs, outs)
    assert n_p1 > 300 and n_p2 > 10 and n_bad >= 3
    assert sum(r["erased"] for r in recs) > 100, "the 4-frame window must have dropped points"
    assert outs[-1][0]["n_map"] > 150


@needs_ref
def test_tracker_oracle_positions_are_the_creating_frames_camera_coordinates():
    """mp_xyz is UnprojectStereo before Rwc / twc (src/frame.cc:171-176): worldpos = Rwc[create_id] * xyz + twc[create_id]."""
    boxes_of = lambda t: BOXES[:1]
    recs = reference_run(7, 4, boxes_of)
    trk = T.Tracker(window=4)
    for t, rec in enumerate(recs):
        trk.reorder(good_names(rec["map_before"]))
        trk.step(rec["cur"]["kps"][:, :2], rec["cur"]["desc"], rec["cur"]["depth_at_kp"], t, boxes=boxes_of(t), F=rec["F"], K4=K4)
    m = recs[-1]["map"]
    mine = {nm: trk.map_xyz[i] for i, nm in enumerate(trk.names())}
    n = 0
    for c, i, b, pos in zip(m["create_id"].tolist(), m["idx"].tolist(), m["bad"].tolist(), m["worldpos"]):
        if b:
            continue
        Tcw = recs[c]["last"]["Tcw"].astype(np.float64)
        Rwc = Tcw[:3, :3].T; twc = -Rwc @ Tcw[:3, 3]
        want = Rwc @ mine[(c, i)].astype(np.float64) + twc
        assert np.allclose(want, pos, rtol=1e-4, atol=1e-3), (c, i, want, pos)
        n += 1
    assert n > 150 and any(not np.allclose(recs[c]["last"]["Tcw"], np.eye(4)) for c in range(1, 4)), "poses must be non-trivial"


@needs_ref
def test_tracker_oracle_without_boxes():
    boxes_of = lambda t: []
    recs = reference_run(6, 5, boxes_of)
    n_bad, n_p1, n_p2 = check(recs, replay(recs, boxes_of))
    assert n_bad == 0 and n_p1 > 200


def fixture_records(path):
    """The committed golden fixture (tests/golden/make_golden_track.py) in the shape run_sequence returns."""
    z = np.load(path)
    recs = []
    for t in range(int(z["n"])):
        g = lambda k: z["%s_%d" % (k, t)]
        kps = np.zeros((len(g("kps")), 6), np.float32); kps[:, :2] = g("kps")
        F = g("F")
        recs.append(dict(
            cur=dict(N=len(kps), kps=kps, desc=g("desc"), depth_at_kp=g("depth"), mp_create_id=g("cur_mp_create"), mp_idx=g("cur_mp_idx"),
                     match_score=g("match_score")),
            map_before=dict(create_id=g("mapb_create_id"), idx=g("mapb_idx"), bad=g("mapb_bad")),
            map=dict(create_id=g("map_create_id"), idx=g("map_idx"), bad=g("map_bad")),
            last=dict(mp_create_id=g("last_mp_create"), mp_idx=g("last_mp_idx")),
            last_after_match=None if t == 0 else dict(mp_bad=g("prev_bad")),
            F=None if F.size == 0 else F, created=int(g("created")), erased=int(g("erased")), boxes=g("boxes")))
    return recs


def test_tracker_oracle_follows_the_golden_fixture_recorded_from_the_reference():
    """Same comparison as above against tests/golden/track_seq5.npz (recorded from the reference's own code here):
    runs where /root/reference does not exist."""
    recs = fixture_records(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "track_seq5.npz"))
    boxes_of = lambda t: recs[t]["boxes"]
    n_bad, n_p1, n_p2 = check(recs, replay(recs, boxes_of))
    assert n_p1 > 300 and n_p2 > 10 and n_bad >= 3 and sum(r["erased"] for r in recs) > 100


def test_own_order_is_survivors_then_new_points():
    """Without an imposed order the map is: survivors in their previous order, then the frame's new points in keypoint order."""
    rng = np.random.default_rng(2)
    trk = T.Tracker(window=2)
    d0 = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    xy = rng.uniform(50, 300, (40, 2)).astype(np.float32)
    z = np.where(np.arange(40) % 3 == 0, -1, 5).astype(np.float32)
    o0 = trk.step(xy, d0, z, 0, K4=K4)
    assert o0["created"] == int((z > 0).sum()) and trk.names() == [(0, i) for i in range(40) if z[i] > 0]
    d1 = d0.copy(); d1[::2] = rng.integers(0, 256, (20, 32), dtype=np.uint8)      # odd rows re-observed exactly
    o1 = trk.step(xy, d1, np.full(40, 4, np.float32), 1, K4=K4)
    re = [i for i in range(40) if i % 2 == 1 and z[i] > 0]
    assert (o1["p1_row_claimed"][re] == 1).all() and (o1["mp_create"][re] == 0).all()
    assert trk.names()[:len(o0["mp_idx"][z > 0])] == [(0, i) for i in range(40) if z[i] > 0]
    o2 = trk.step(xy, d1, np.full(40, 4, np.float32), 2, K4=K4)                    # window 2: frame-0 points age out of the map
    assert all(c >= 1 for c, _ in trk.names())
    assert (o2["mp_create"][re] == 0).all(), "a point that left the map is still tracked frame to frame (pass 1)"


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_tracker_state_invariants_on_random_frames(seed):
    """Random frames with controlled re-observations, boxes and a small capacity: after every frame the state is
    consistent — links and map rows point at each other, a keypoint's frozen descriptor is the descriptor its point had
    at creation, point names are unique, nothing older than the window is in the map, the capacity holds."""
    rng = np.random.default_rng(seed)
    cap, window = 260, 3
    trk = T.Tracker(window=window, map_cap=cap)
    born = {}                                   # (create_id, idx) -> descriptor at creation
    prev = None
    F = np.array([[0, 0, 0], [0, 0, -1.0], [0, 1.0, 0]])       # epipolar lines are the rows: distance = |dy|
    for t in range(9):
        n = int(rng.integers(120, 180))
        desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if prev is not None:                    # re-observe about half of the last frame with 0-3 flipped bits
            k = min(len(prev), n) // 2
            src = rng.permutation(len(prev))[:k]
            flips = np.packbits(rng.random((k, 256)) < 0.006, axis=1, bitorder="little")
            desc[rng.permutation(n)[:k]] = prev[src] ^ flips
        xy = np.stack([rng.uniform(0, 400, n), rng.uniform(0, 240, n)], 1).astype(np.float32)
        depth = np.where(rng.random(n) < 0.8, rng.uniform(1, 50, n), -1).astype(np.float32)
        boxes = [[100, 220, 50, 150]] if t % 2 else None
        o = trk.step(xy, desc, depth, t, boxes=boxes, F=F if t % 2 else None, K4=K4)
        for j in np.nonzero(o["mp_create"] == t)[0]:
            born[(t, int(j))] = desc[j].copy()
        # names are unique and every one was born with the descriptor the state holds
        names = trk.names()
        assert len(set(names)) == len(names) <= cap
        for (c, i), d in zip(names, trk.map_desc):
            assert c > t - window and (born[(c, i)] == d).all()
        # links: map row r <-> keypoint map_link[r]
        for r, j in enumerate(trk.map_link):
            if j >= 0:
                assert trk.prev_map_row[j] == r and trk.prev_live[j] == 1
        for j, r in enumerate(trk.prev_map_row):
            if r >= 0:
                assert trk.map_link[r] == j
                assert (trk.prev_desc[j] == trk.map_desc[r]).all() and (trk.prev_create[j], trk.prev_idx[j]) == names[r]
        # every live keypoint carries the frozen descriptor of its point; the frame's own descriptors are kept beside it
        lv = trk.prev_live == 1
        for j in np.nonzero(lv)[0]:
            assert (trk.prev_desc[j] == born[(int(trk.prev_create[j]), int(trk.prev_idx[j]))]).all()
        assert (trk.last_desc == desc).all() and (trk.prev_create[~lv] == -1).all()
        # no new point inside a box grown by 5 px
        if boxes:
            b = boxes[0]
            inside = (xy[:, 0] > b[0] - 5) & (xy[:, 0] < b[1] + 5) & (xy[:, 1] > b[2] - 5) & (xy[:, 1] < b[3] + 5)
            assert not (o["mp_create"][inside] == t).any()
        prev = desc
    assert len(born) > 400
