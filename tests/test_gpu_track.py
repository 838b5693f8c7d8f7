"""GPU parity of the device-resident tracker state (csrc/track.cu, svo_track_* and svo_frame_in.track_seq) against
oracle/track.py, the CPU restatement of what Tracking::Track carries from frame to frame (src/Tracking.cc:225-250),
which tests/test_oracle_track.py pins to the reference's own code.

Every frame goes through svo_batch_submit with nothing but images (+ boxes / F); the oracle is stepped with the
keypoints, descriptors and stereo depths the same batch returned (the extractor and the stereo stage have their own
parity tests) and everything downstream must be identical: BF over the last frame's OWN descriptors, pass 1 over the
FROZEN map-point descriptors with the veto, pass 2 over the local map, the point every keypoint owns afterwards, and
the whole next state (bit for bit, positions included)."""
import numpy as np
import pytest

import synth
from oracle import track as T

pytestmark = pytest.mark.gpu

CAL = synth.KITTI_04_12
K4 = (CAL["fx"], CAL["fy"], CAL["cx"], CAL["cy"])
BF = float(CAL["bf"]); BASE = float(CAL["bf"] / CAL["fx"])
F_TEST = np.array([[1.1e-9, 2.3e-7, -3.1e-4], [-2.2e-7, 0.9e-9, 0.8312], [2.9e-4, -0.8297, 1.0]], np.float64)
BOXES = np.array([[300, 700, 100, 300], [900, 1100, 50, 200], [20, 180, 200, 360]], np.int32)
SHAPE = (240, 400)


@pytest.fixture(scope="module")
def svo():
    import svo as S
    return S


def frame_dict(img, seq, t, boxes=None, F=None):
    d = dict(left=img[0], right=img[1], bf=BF, baseline=BASE, track_seq=seq, frame_id=t, K=K4)
    if boxes is not None:
        d["boxes"] = boxes
    if F is not None:
        d["F"] = F
    return d


def compare_frame(r, o, tag):
    assert r["status"] == 0, tag
    assert r["n_prev"] == o["n_prev"] and r["n_map"] == o["n_map"], (tag, r["n_prev"], o["n_prev"], r["n_map"], o["n_map"])
    keys = ["claim_row", "mp_create"]
    if o["n_prev"]:
        keys += ["bf_idx", "bf_dist", "bf_keep", "p1_best_idx", "p1_best", "p1_second", "p1_row_claimed"]
        if r.get("p1_row_bad") is not None and len(r["p1_row_bad"]):
            keys.append("p1_row_bad")
    if o["n_map"]:
        keys.append("p2_row_claimed")
    for k in keys:
        assert np.array_equal(r[k], o[k]), (tag, k, int((np.asarray(r[k]) != np.asarray(o[k])).sum()))
    assert (r["mp_xyz"].view(np.uint32) == o["mp_xyz"].view(np.uint32)).all(), (tag, "mp_xyz")


def compare_state(st, trk, tag):
    assert st["n_prev"] == len(trk.prev_desc) and st["n_map"] == len(trk.map_desc), tag
    for k in ("last_desc", "prev_desc", "prev_live", "prev_map_row", "prev_create", "map_desc", "map_create", "map_link"):
        assert np.array_equal(st[k], getattr(trk, k)), (tag, k)
    for k in ("prev_xyz", "prev_xy", "map_xyz"):
        assert (st[k].view(np.uint32) == getattr(trk, k).view(np.uint32)).all(), (tag, k)


def run_sequences(ctx, seeds, n_frames, lanes_of, boxes_of, shape=SHAPE, trackers=None, check_state=True):
    seqs = [synth.Sequence(shape, seed=s) for s in seeds]
    trackers = trackers or [T.Tracker(window=4, map_cap=ctx.track_cap) for _ in seeds]
    stats = dict(p1=0, p2=0, bad=0, aged=0)
    for t in range(n_frames):
        lane = lanes_of(t)
        frames = []
        for s, sq in enumerate(seqs):
            bx, F = boxes_of(s, t)
            frames.append(frame_dict(sq.frame(t), s, t, bx, F))
        ctx.batch_submit(lane, frames)
        ctx.batch_wait(lane)
        for s in range(len(seqs)):
            r = ctx.batch_result(lane, s)
            bx, F = boxes_of(s, t)
            xy = np.stack([r["kp_left"]["x"], r["kp_left"]["y"]], 1)
            before = len(trackers[s].map_desc)
            o = trackers[s].step(xy, r["desc_left"], r["depth"], t, boxes=bx, F=F, K4=K4)
            compare_frame(r, o, "seq %d frame %d" % (s, t))
            if o["n_prev"]:
                stats["p1"] += int(o["p1_row_claimed"].sum()); stats["bad"] += int(o["p1_row_bad"].sum())
            if o["n_map"]:
                stats["p2"] += int(o["p2_row_claimed"].sum())
            stats["aged"] += max(0, before + o["created"] - len(trackers[s].map_desc))
            if check_state:
                compare_state(ctx.track_state(s), trackers[s], "seq %d after frame %d" % (s, t))
    return trackers, stats


def test_tracked_sequences_match_the_oracle(svo):
    """Three sequences, seven frames, alternating lanes (so every batch waits for the state the other lane's batch left:
    the split-graph path), boxes on some frames with and without F."""
    ctx = svo.Context(SHAPE[1], SHAPE[0], nfeatures=500, max_batch=3, lanes=2, max_rows=3000)
    ctx.track_create(3, 3000, 4)
    for s in range(3):
        ctx.track_reset(s)

    def boxes_of(s, t):
        sc = np.array([0.32, 0.32, 0.64, 0.64])        # the K-shape boxes scaled to 400 x 240
        bx = (BOXES * sc).astype(np.int32)
        if s == 0:
            return (bx, F_TEST) if t % 2 == 1 else (bx[:1], F_TEST)
        if s == 1:
            return (bx[:2], None) if t >= 2 else (None, None)       # boxes without F: no veto, but no points inside them
        return None, None

    _, st = run_sequences(ctx, (11, 12, 13), 7, lambda t: t % 2, boxes_of)
    assert st["p1"] > 500 and st["p2"] > 20 and st["aged"] > 300, st
    assert st["bad"] >= 1, "the veto should have marked some points bad"
    ctx.close()


def test_tracked_one_lane_with_ballast_and_a_full_map(svo):
    """One lane (no cross-lane wait: the single-graph path), a map seeded with ballast rows that never age out, and a
    capacity small enough that new points stop fitting: they live in the frame only."""
    rng = np.random.default_rng(3)
    ballast = rng.integers(0, 256, (700, 32), dtype=np.uint8)
    ctx = svo.Context(SHAPE[1], SHAPE[0], nfeatures=500, max_batch=2, lanes=1, max_rows=820)
    ctx.track_create(2, 820, 4)
    ctx.track_reset(0, ballast)
    ctx.track_reset(1)
    trackers = [T.Tracker(window=4, ballast=ballast, map_cap=820), T.Tracker(window=4, map_cap=820)]
    trk, st = run_sequences(ctx, (21, 22), 6, lambda t: 0, lambda s, t: (None, None), trackers=trackers)
    assert len(trk[0].map_desc) == 820 and (trk[0].map_create[:700] == T.BALLAST).all(), "ballast stays, the map is full"
    assert (trk[0].prev_live == 1).sum() > (trk[0].prev_map_row >= 0).sum(), "some points live in the frame only"
    # reset: the sequence starts over
    ctx.track_reset(1)
    st1 = ctx.track_state(1)
    assert st1["n_prev"] == 0 and st1["n_map"] == 0
    ctx.close()


def test_tracked_and_untracked_frames_share_a_batch(svo):
    """A tracked frame next to a frame that brings its own previous-frame descriptors and map (the untracked contract)."""
    from oracle import oracle as O
    ctx = svo.Context(SHAPE[1], SHAPE[0], nfeatures=500, max_batch=2, lanes=1, max_rows=2000)
    ctx.track_create(1, 2000, 4)
    ctx.track_reset(0)
    seq = synth.Sequence(SHAPE, seed=31); other = synth.Sequence(SHAPE, seed=32)
    trk = T.Tracker(window=4, map_cap=2000)
    rng = np.random.default_rng(5)
    prev = None
    for t in range(3):
        fr = [frame_dict(seq.frame(t), 0, t), dict(left=other.frame(t)[0], right=other.frame(t)[1], bf=BF, baseline=BASE)]
        if prev is not None:
            mp = rng.integers(0, 256, (900, 32), dtype=np.uint8); mp[:300] = prev[:300]
            fr[1].update(prev_desc=prev, map_desc=mp)
        ctx.batch_submit(0, fr); ctx.batch_wait(0)
        r0, r1 = ctx.batch_result(0, 0), ctx.batch_result(0, 1)
        xy = np.stack([r0["kp_left"]["x"], r0["kp_left"]["y"]], 1)
        compare_frame(r0, trk.step(xy, r0["desc_left"], r0["depth"], t, K4=K4), "tracked frame %d" % t)
        if prev is not None:
            bi, bd, bk = O.match_bf(r1["desc_left"], prev)
            assert np.array_equal(r1["bf_idx"], bi) and np.array_equal(r1["bf_dist"], bd) and np.array_equal(r1["bf_keep"], bk)
            p1 = O.match_greedy(prev, r1["desc_left"], 0)
            assert np.array_equal(r1["p1_row_claimed"], p1["row_claimed"])
            p2 = O.match_greedy(mp, r1["desc_left"], 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_base=len(prev))
            assert np.array_equal(r1["claim_row"], p2["claim_row"])
            assert "mp_create" not in r1
        prev = r1["desc_left"].copy()
    ctx.close()


def test_track_argument_errors(svo):
    ctx = svo.Context(SHAPE[1], SHAPE[0], nfeatures=300, max_batch=2, lanes=1, max_rows=1000)
    img = synth.Sequence(SHAPE, seed=1).frame(0)
    with pytest.raises(svo.SvoError):
        ctx.batch_submit(0, [frame_dict(img, 0, 0)])              # no tracker states yet
    with pytest.raises(svo.SvoError):
        ctx.track_create(1, 5000, 4)                              # map capacity above max_rows
    ctx.track_create(2, 1000, 4)
    with pytest.raises(svo.SvoError):
        ctx.track_create(2, 1000, 4)                              # only once
    with pytest.raises(svo.SvoError):
        ctx.batch_submit(0, [frame_dict(img, 0, 0), frame_dict(img, 0, 1)])   # one sequence twice in a batch
    with pytest.raises(svo.SvoError):
        ctx.batch_submit(0, [frame_dict(img, 2, 0)])              # sequence out of range
    with pytest.raises(svo.SvoError):
        ctx.track_reset(5)
    ctx.batch_submit(0, [frame_dict(img, 1, 0)]); ctx.batch_wait(0)
    assert ctx.batch_result(0, 0)["n_prev"] == 0
    ctx.close()


def test_tracked_kitti_shape_2000_features(svo):
    """The bench's own shape: 1241x376, 2000 features, 5000-row capacity, two sequences over five frames."""
    ctx = svo.Context(1241, 376, nfeatures=2000, max_batch=2, lanes=2, max_rows=5000, skip_match_score=False)
    ctx.track_create(2, 5000, 4)
    ctx.track_reset(0); ctx.track_reset(1)
    bx = BOXES
    _, st = run_sequences(ctx, (41, 42), 5, lambda t: t % 2, lambda s, t: (bx, F_TEST) if s == 0 else (None, None),
                          shape=synth.K_SHAPE, check_state=False)
    assert st["p1"] > 2000 and st["p2"] > 50, st
    ctx.close()


@pytest.mark.parametrize("rows", [None, "100"])
def test_compact_and_left_only_outputs_are_the_same_results(svo, rows, monkeypatch):
    """svo_set_outputs: the compact strided copies (SVO_OUT_COMPACT) and leaving the right image's features on the device
    (SVO_OUT_NO_RIGHT) change what crosses PCIe, not the results.  With SVO_B200_COMPACT_ROWS=100 every frame holds more
    rows than the copies carry, so svo_batch_wait's tail fetch runs for every array."""
    if rows:
        monkeypatch.setenv("SVO_B200_COMPACT_ROWS", rows)
    ctx = svo.Context(SHAPE[1], SHAPE[0], nfeatures=500, max_batch=2, lanes=1, max_rows=2000)
    ctx.track_create(2, 2000, 4)
    seqs = [synth.Sequence(SHAPE, seed=s) for s in (51, 52)]
    rng = np.random.default_rng(9)
    results = {}
    base_desc = []      # frame t's left descriptors of the untracked sequence (the pose-inputs mode does not return them)
    for flags in (0, svo.OUT_COMPACT, svo.OUT_COMPACT | svo.OUT_NO_RIGHT, svo.OUT_COMPACT | svo.OUT_POSE_INPUTS):
        ctx.set_outputs(flags)
        ctx.track_reset(0); ctx.track_reset(1)
        out = []
        prev = None
        for t in range(4):
            # frame 0 of the batch is tracked (with a veto), frame 1 brings 1500 pass-1 rows of its own (> the compact width)
            fr = [frame_dict(seqs[0].frame(t), 0, t, (BOXES * np.array([0.32, 0.32, 0.64, 0.64])).astype(np.int32), F_TEST),
                  dict(left=seqs[1].frame(t)[0], right=seqs[1].frame(t)[1], bf=BF, baseline=BASE)]
            if prev is not None:
                big = np.concatenate([prev, np.random.default_rng(t).integers(0, 256, (1500 - len(prev), 32), dtype=np.uint8)], 0)
                fr[1].update(prev_desc=big, map_desc=prev[::2].copy())
            ctx.batch_submit(0, fr); ctx.batch_wait(0)
            out.append([ctx.batch_result(0, 0), ctx.batch_result(0, 1)])
            prev = base_desc[t] if flags else out[-1][1]["desc_left"].copy()
            if not flags:
                base_desc.append(prev)
        results[flags] = out
    ctx.set_outputs(0)
    base = results[0]
    for flags, out in results.items():
        for t in range(4):
            for i in range(2):
                a, b = base[t][i], out[t][i]
                for k, v in a.items():
                    if k in ("kp_right", "desc_right") and flags & (svo.OUT_NO_RIGHT | svo.OUT_POSE_INPUTS):
                        assert b[k] is None
                        continue
                    if flags & svo.OUT_POSE_INPUTS and k not in ("status", "n_left", "n_right", "n_stereo", "kp_left", "depth", "claim_row",
                                                                  "n_prev", "n_map", "mp_create", "mp_xyz"):
                        assert b.get(k) is None, k          # stays on the device
                        continue
                    if isinstance(v, np.ndarray):
                        assert v.dtype == b[k].dtype and v.shape == b[k].shape and v.tobytes() == b[k].tobytes(), (flags, t, i, k)
                    else:
                        assert v == b[k], (flags, t, i, k)
    assert base[3][1]["n_prev"] == 1500 and base[3][0]["n_prev"] > 300
    ctx.close()
