"""Adversarial feature sets through the reference's OWN poseEstimationPnP (src/pnpmatch.cc:33-251, compiled unmodified
into oracle/_ref/libsvo_ref.so): hand-made descriptors answer its cv::ORB calls (oracle/ref.py:ORB_OVERRIDE), so the cases
image sequences never produce — long chains of identical descriptors, exact ties, distances sitting on the thresholds
(best = 14 / 15 in pass 1; best = 29 / 30 and second / best = 2.0 / 2.05 in pass 2), best = 0 (match_score = inf),
vetoed claims that free a column for a later row — are decided by the reference itself (tests/adversarial_sets.py) and
compared with oracle/svo_matchers.c here and, through the recorded fixture tests/golden/ref_adversarial.npz
(tests/golden/make_golden_ref_adversarial.py), with svo_match_greedy on the GPU box."""
import os

import numpy as np
import pytest

import adversarial_sets as A
import synth
from oracle import ref as R

from test_ref_pin import check_run

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
BF = np.float32(CAL["bf"])
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ref_adversarial.npz")


def check_case(seed, boxes, run, matcher=None):
    """The reference's decisions of one case against an implementation of the scans (default: the oracle)."""
    if seed == 2:                                # a box over the whole image: no map points, no point pairs for F
        assert run["created"] == 0 and run["F"]["F"] is None
        run["F"]["F"] = np.zeros((3, 3))         # never read: there is no live row
    p1, p2 = check_run(run, boxes, matcher)
    if seed == 2:
        assert p1["row_claimed"].sum() == 0 and p2["row_claimed"].sum() == 0
    elif not boxes:
        best, claimed = p1["best"], p1["row_claimed"].astype(bool)
        assert claimed[:20].all(), "the chain of identical rows claims twenty columns one after the other"
        assert (best[:12] == 0).all() and (best[12:17] == 3).all() and (best[17:20] == 14).all()
        assert np.isinf(run["cur"]["match_score"][:11]).all()           # second / 0
        assert (best[claimed] < 15).all() and claimed[41:60:2].all() and not claimed[40:60:2].any()
        assert p2["row_claimed"].sum() >= 15                            # 29 and 41 / 20 claim, 30 and 40 / 20 do not
    else:
        assert p1["row_bad"].sum() >= 5, "the box and the vertical offsets should veto some would-be claims"


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (/root/reference absent)")
@pytest.mark.parametrize("seed", sorted(A.CASES))
def test_oracle_follows_the_reference_on_adversarial_sets(seed):
    boxes = A.CASES[seed]
    check_case(seed, boxes, A.run_reference(seed, boxes, K, BF))


@pytest.mark.parametrize("seed", sorted(A.CASES))
def test_oracle_follows_the_recorded_reference_decisions(seed):
    """The same through the committed fixture (what the GPU box has)."""
    check_case(seed, A.CASES[seed], A.unpack_run(np.load(GOLDEN), "c%d." % seed))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (/root/reference absent)")
def test_fixture_is_what_the_reference_computes_now():
    g = np.load(GOLDEN)
    for seed, boxes in A.CASES.items():
        for k, v in A.pack_run(A.run_reference(seed, boxes, K, BF), "c%d." % seed).items():
            if ".before.map." in k:              # the std::set's pointer order differs from run to run (src/pnpmatch.cc:160)
                rows = lambda a: sorted(bytes(np.ascontiguousarray(r)) for r in np.asarray(a))
                assert rows(v) == rows(g[k]), k
            elif k.endswith("cur.mp_idx") or k.endswith("cur.mp_create_id"):
                continue                         # pass-2 claims depend on that order; check_case verifies them per run
            else:
                assert np.array_equal(np.asarray(v), g[k], equal_nan=True), k


@pytest.mark.gpu
@pytest.mark.parametrize("seed", sorted(A.CASES))
def test_gpu_matchers_follow_the_recorded_reference_decisions(seed):
    """svo_match_greedy (pass 1 with the veto, then pass 2 in the reference's set order) against the decisions the
    reference's own code took on the adversarial sets: match_score bits, mp->bad, the final CurrentFrame->MapPoints."""
    import svo
    ctx = svo.Context(1241, 376, nfeatures=2000, max_batch=1, lanes=1, max_rows=5000)
    try:
        check_case(seed, A.CASES[seed], A.unpack_run(np.load(GOLDEN), "c%d." % seed), matcher=ctx.match_greedy)
    finally:
        ctx.close()


@pytest.mark.parametrize("seed", sorted(A.CASES))
def test_bf_matcher_on_adversarial_sets_is_cv2s(seed):
    """cv::BFMatcher(NORM_HAMMING)::match as src/pnpmatch.cc:278 calls it, on the hand-made sets: a query with twenty
    identical train rows at distance 0 takes the first; the filter threshold is max(2 * 0, 30) = 30 (:291-299)."""
    import cv2
    from oracle import oracle as O
    (_, d_last), (_, d_cur) = A.feature_sets(seed)
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(d_cur, d_last)
    idx, dist, keep = O.match_bf(d_cur, d_last)
    assert [x.queryIdx for x in m] == list(range(len(d_cur)))
    assert (np.array([x.trainIdx for x in m]) == idx).all() and (np.array([x.distance for x in m]) == dist).all()
    assert dist.min() == 0 and (keep == (dist <= 30)).all() and 0 < keep.sum() < len(keep)
    assert (idx[(d_cur == d_last[0]).all(1)] == 0).all()           # the chain's columns all name the FIRST identical row


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (/root/reference absent)")
@pytest.mark.parametrize("seed", [0, 3])
def test_F_inputs_on_adversarial_sets(seed):
    """poseEstimation2D_2D (src/pnpmatch.cc:302-337) on the hand-made sets: the point pairs the reference hands to
    findFundamentalMat are the oracle's filtered BF matches minus current keypoints inside a box grown by 10 px."""
    from oracle import oracle as O
    boxes = A.CASES[seed]
    run = A.run_reference(seed, boxes, K, BF)
    cur, last = run["cur"], run["last"]
    idx, dist, keep = O.match_bf(cur["desc"], last["desc"])
    q = np.nonzero(keep)[0]
    x, y = cur["kps"][q, 0], cur["kps"][q, 1]
    inside = np.zeros(len(q), bool)
    for b in boxes:
        inside |= (x > b[0] - 10) & (x < b[1] + 10) & (y > b[2] - 10) & (y < b[3] + 10)
    q = q[~inside]
    assert np.array_equal(cur["kps"][q, :2], run["F"]["p1"]) and np.array_equal(last["kps"][idx[q], :2], run["F"]["p2"])
    assert len(q) >= 8 and (not boxes or inside.any())


# ---- the tracker oracle (oracle/track.py) over hand-made SEQUENCES through the reference's frame loop -----------------

def _track_check(recs, boxes_of):
    from test_oracle_track import check, replay
    return check(recs, replay(recs, boxes_of))


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (/root/reference absent)")
@pytest.mark.parametrize("seed,with_boxes", [(0, False), (1, True)])
def test_tracker_oracle_on_hand_made_sequences(seed, with_boxes):
    """Seven frames of hand-made features (tests/adversarial_sets.py:sequence_sets) through the reference's frame loop:
    re-observed points, points coming back after two and three frames (pass 2), a block of identical descriptors in every
    frame (claim chains in both passes), the 4-frame window — every frame's MapPoints, match_score, bad flags, created
    points and local map equal the tracker oracle's."""
    boxes_of = (lambda t: [[200, 700, 60, 250]] if t % 2 else [[900, 1100, 0, 200]]) if with_boxes else (lambda t: [])
    recs = A.run_reference_sequence(seed, 7, boxes_of, K, BF)
    n_bad, n_p1, n_p2 = _track_check(recs, boxes_of)
    assert n_p1 > 600 and n_p2 > 60, (n_p1, n_p2)
    assert (n_bad > 0) == with_boxes
    assert sum(r["erased"] for r in recs) > 100


def test_tracker_oracle_on_a_hand_made_sequence_through_tracking_track(tmp_path):
    """The same through Tracking::Track itself (src/Tracking.cc compiled unmodified, oracle/_ref/libsvo_ref_g2o.so):
    Tracking::init, poseEstimationPnP, Optimizer::PoseOptimization, createmappoint and the window as main.cpp drives them."""
    from oracle import ref_g2o as RG
    if not RG.available():
        pytest.skip("/root/reference (or a prebuilt oracle/_ref/libsvo_ref_g2o.so) is not present")
    boxes_of = lambda t: [[200, 700, 60, 250]] if t % 2 else []
    recs = A.run_reference_sequence(2, 6, boxes_of, K, BF, runner=RG.run_tracking, tmpdir=tmp_path)
    n_bad, n_p1, n_p2 = _track_check(recs, boxes_of)
    assert n_p1 > 500 and n_p2 > 40 and n_bad > 0, (n_bad, n_p1, n_p2)


@pytest.mark.skipif(not R.available(), reason="oracle/_ref cannot be built here (/root/reference absent)")
def test_a_slice_of_the_fuzz():
    """tools/fuzz_ref_pin.py: random hand-made sets (duplicate blocks, distances around the thresholds, rival columns for the
    ratio test, random boxes and positions, keypoints without depth) through the reference's poseEstimationPnP against the
    oracle; 440 seeds were run when this was written (none diverged), eight of them here."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_ref_pin as Z
    tot = np.zeros(3, np.int64)
    for seed in range(8):
        tot += Z.run(seed)
    assert tot[0] > 500 and tot[1] > 50 and tot[2] > 200, tot
