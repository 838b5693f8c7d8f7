"""Pins the CPU oracle to the REFERENCE'S OWN CODE: oracle/_ref/libsvo_ref.so is /root/reference/src/pnpmatch.cc,
src/frame.cc and src/mappoint.cc compiled unmodified (oracle/Makefile `ref`, oracle/ref_stubs/minicv.hpp), with its
OpenCV calls answered by the real cv2.  The reference is driven as Tracking::Track / Tracklastframe drive it
(src/Tracking.cc:184-250, :114) and every result on the hot path is compared with oracle/svo_oracle.c:
DescriptorDistance, pass 1 with the YOLO-box / epipolar veto, pass 2 in the std::set's own order, match_score,
computekeypoint_r, disp2Depth, createmappoint and UnprojectStereo."""
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="/root/reference (or a prebuilt oracle/_ref) is not present")

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
BF = np.float32(CAL["bf"])
BOXES = [[300, 700, 100, 300], [900, 1100, 50, 200], [20, 180, 200, 360]]


def run_reference(seed, boxes, shape=synth.K_SHAPE):
    """Two frames through the reference exactly as Tracking::Track does; returns everything the tests compare."""
    seq = synth.Sequence(shape, seed=seed)
    frames = (seq.frame(0), seq.frame(1))
    disps = (synth.dense_disparity(shape, 2 * seed), synth.dense_disparity(shape, 2 * seed + 1))
    return R.run_two_frames(frames, disps, K, BF, boxes)


@pytest.fixture(scope="module")
def run5():
    return run_reference(5, BOXES)


def test_descriptor_distance_is_the_oracles():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (300, 32), dtype=np.uint8); b = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    b[:50] = a[:50] ^ np.packbits(rng.random((50, 256)) < 0.03, axis=1)
    for x, y in zip(a, b):
        assert R.descriptor_distance(x, y) == O.hamming(x, y)
    assert R.descriptor_distance(a[0], a[0]) == 0
    assert R.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def replay_with_oracle(run, boxes, matcher=None):
    """The oracle's pass 1 (+ veto) and pass 2 on the inputs the reference saw; pass 2 in the set's own order.
    matcher: another implementation with O.match_greedy's signature (the GPU tests pass svo.Context.match_greedy)."""
    match_greedy = matcher or O.match_greedy
    last0, map0 = run["before"]["last"], run["before"]["map"]
    cur, last = run["cur"], run["last"]
    M = last0["N"]
    live = (last0["mp_create_id"] >= 0).astype(np.uint8)[:M]
    rows = np.zeros((M, 32), np.uint8)
    rows[:len(last0["desc"])] = last0["desc"][:M]     # m_descriptor is the creating frame's row, frozen (src/mappoint.cc:12)
    # find_feature_matches overwrote both frames' keypoints_l with a second ORB pass (src/pnpmatch.cc:306): the veto
    # reads those (they are the same keypoints: detect+compute == detectAndCompute)
    veto = dict(boxes=boxes, F=run["F"]["F"], row_xy=last["kps"][:M, :2], cur_xy=cur["kps"][:, :2]) if len(boxes) else None
    p1 = match_greedy(rows, cur["desc"], 0, row_live=live, veto=veto)
    order = [(int(c), int(i)) for c, i in zip(map0["create_id"], map0["idx"])]
    assert all(c == 0 for c, _ in order)
    idx2 = np.array([i for _, i in order], np.int64)
    live2 = np.ones(len(idx2), np.uint8)
    live2[p1["row_bad"][idx2] == 1] = 0               # mp->bad (src/pnpmatch.cc:163)
    live2[p1["row_claimed"][idx2] == 1] = 0           # observations.count(CurrentFrame) (:165)
    p2 = match_greedy(map0["desc"], cur["desc"], 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                      row_base=10 ** 6)
    return live, p1, idx2, p2


def check_run(run, boxes, matcher=None):
    cur, last = run["cur"], run["last"]
    live, p1, idx2, p2 = replay_with_oracle(run, boxes, matcher)
    M = len(live)
    # pass 1: match_score (src/pnpmatch.cc:99), bad flags (:141), claims (:151)
    lv = live.astype(bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        score = p1["second"].astype(np.float32) / p1["best"].astype(np.float32)
    assert (cur["match_score"][:M][lv].view(np.uint32) == score[lv].view(np.uint32)).all()
    assert (cur["match_score"][:M][~lv] == -1).all()
    assert (last["mp_bad"][:M][lv] == p1["row_bad"][lv]).all()
    # final CurrentFrame->MapPoints: pass-1 claims name (0, i); pass-2 claims name the set's row
    expect = np.full(cur["N"], -1, np.int64)
    for j in np.nonzero(p2["claim_row"] >= 0)[0]:
        r = int(p2["claim_row"][j])
        expect[j] = r if r < 10 ** 6 else idx2[r - 10 ** 6]
    got = np.where(cur["mp_create_id"] == 0, cur["mp_idx"], -1)
    assert (got == expect).all()
    return p1, p2


def test_pose_estimation_pnp_passes_match_the_oracle(run5):
    p1, p2 = check_run(run5, BOXES)
    assert p1["row_claimed"].sum() > 40 and p2["row_claimed"].sum() > 3
    assert p1["row_bad"].sum() >= 3, "the boxes should veto some would-be matches"
    # some vetoed rows would have claimed a column that a LATER row then takes: the veto changes later decisions
    no_veto = O.match_greedy(run5["before"]["last"]["desc"], run5["cur"]["desc"], 0,
                             row_live=(run5["before"]["last"]["mp_create_id"] >= 0).astype(np.uint8))
    assert no_veto["row_claimed"].sum() == p1["row_claimed"].sum() + p1["row_bad"].sum() or \
        (no_veto["claim_row"] != p1["claim_row"]).any()


@pytest.mark.parametrize("seed,boxes", [(6, []), (7, [[640, 1241, 0, 376]]), (8, BOXES[:1])])
def test_pose_estimation_pnp_other_inputs(seed, boxes):
    """No boxes (veto off); one box over the right half of the image (no map points are created inside it,
    src/frame.cc:196-207, so only matches that cross its border reach the epipolar test); another sequence."""
    run = run_reference(seed, boxes)
    p1, p2 = check_run(run, boxes)
    if not boxes:
        assert p1["row_bad"].sum() == 0


def test_find_feature_matches_filter_and_F_inputs(run5):
    """poseEstimation2D_2D (src/pnpmatch.cc:302-337): BF matches of the re-extracted descriptors, filtered by
    d <= max(2 * min, 30), minus current keypoints inside a box grown by 10 px, are what findFundamentalMat gets."""
    cur, last = run5["cur"], run5["last"]
    idx, dist, keep = O.match_bf(cur["desc"], last["desc"])
    p1, p2 = [], []
    for q in np.nonzero(keep)[0]:
        x, y = cur["kps"][q, 0], cur["kps"][q, 1]
        if any(x > b[0] - 10 and x < b[1] + 10 and y > b[2] - 10 and y < b[3] + 10 for b in BOXES):
            continue
        p1.append(cur["kps"][q, :2]); p2.append(last["kps"][idx[q], :2])
    assert np.array_equal(np.array(p1, np.float32), run5["F"]["p1"]) and np.array_equal(np.array(p2, np.float32), run5["F"]["p2"])


def test_stereo_fields_and_map_points(run5):
    """computekeypoint_r (src/frame.cc:122-138), disp2Depth (:140-164), createmappoint (:182-238), UnprojectStereo (:166-180)."""
    s0, last0, map0 = run5["f0"], run5["before"]["last"], run5["before"]["map"]
    disp, depth = run5["disp0"], run5["depth0"]
    # disp2Depth over the whole image is the oracle's
    assert (depth.view(np.uint32) == O.disp2depth(disp, BF).view(np.uint32)).all()
    # keypoints_r: rx sticks at its last value where the disparity is -1
    kx, ky = s0["kps"][:, 0], s0["kps"][:, 1]
    rx = np.float32(-1); exp = np.empty(len(kx), np.float32)
    for i in range(len(kx)):
        d = disp[int(ky[i]), int(kx[i])]
        if d != -1:
            rx = kx[i] - d
        exp[i] = rx
    assert (s0["keypoints_r"][:, 0].view(np.uint32) == exp.view(np.uint32)).all() and (s0["keypoints_r"][:, 1] == ky).all()
    assert (exp == -1).any() or (disp[ky.astype(int), kx.astype(int)] == -1).any()
    # createmappoint: a point for every keypoint with depth > 0 outside every box grown by 5 px
    z = depth[ky.astype(np.int64), kx.astype(np.int64)]
    want = z > 0
    for b in BOXES:
        want &= ~((kx > b[0] - 5) & (kx < b[1] + 5) & (ky > b[2] - 5) & (ky < b[3] + 5))
    assert ((last0["mp_create_id"] >= 0) == want[:last0["N"]]).all() and run5["created"] == int(want.sum())
    # worldpos = Rwc * ((u-cx) z (1/fx), (v-cy) z (1/fy), z) + twc with the identity pose
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    for c, i, pos in zip(map0["create_id"], map0["idx"], map0["worldpos"]):
        u, v, zz = kx[i], ky[i], z[i]
        x = (u - cx) * zz * (np.float32(1) / fx); y = (v - cy) * zz * (np.float32(1) / fy)
        assert (pos.view(np.uint32) == np.array([x, y, zz], np.float32).view(np.uint32)).all()
    assert (map0["desc"] == last0["desc"][map0["idx"]]).all()


def test_unproject_stereo_with_a_pose_follows_opencv_gemm():
    """UnprojectStereo under a non-trivial pose: x3D = Rwc * x3Dc + twc.  minicv's small-matrix product (float
    products summed left to right, then the add) is checked against the real cv2.gemm on the same operands."""
    import cv2
    rng = np.random.default_rng(3)
    T = np.eye(4, dtype=np.float32)
    T[:3, :3] = synth.rodrigues([0.03, -0.2, 0.1]).astype(np.float32); T[:3, 3] = [0.4, -0.1, 2.0]
    seq = synth.Sequence(seed=1)
    L, Rr = seq.frame(0)
    f = R.Frame(L, Rr, K, BF, [], 0.0, 0)
    f.set_pose(T)
    Rwc = np.ascontiguousarray(T[:3, :3].T); tcw = T[:3, 3:4].copy()
    twc = cv2.gemm(Rwc, tcw, -1.0, None, 0.0)                       # twc = -Rwc*tcw (src/frame.cc:72)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    for _ in range(200):
        u, v, z = np.float32(rng.uniform(0, 1241)), np.float32(rng.uniform(0, 376)), np.float32(rng.uniform(0.5, 80))
        xc = np.array([[(u - cx) * z * (np.float32(1) / fx)], [(v - cy) * z * (np.float32(1) / fy)], [z]], np.float32)
        want = cv2.gemm(Rwc, xc, 1.0, twc, 1.0)[:, 0]               # what the MatExpr Rwc*x3Dc+twc evaluates to
        got = f.unproject(u, v, z)
        assert (got.view(np.uint32) == want.view(np.uint32)).all()
    assert f.unproject(10, 10, 0.0) is None and f.unproject(10, 10, -1.0) is None


def test_minicv_bfmatcher_is_cv2s():
    """minicv.hpp's DescriptorMatcher stands in for cv::BFMatcher inside libsvo_ref.so: same first-minimum rule."""
    import cv2
    rng = np.random.default_rng(4)
    t = rng.integers(0, 256, (400, 32), dtype=np.uint8); q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t[100:140] = t[7]; q[:50] = t[rng.integers(0, 400, 50)]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
    idx, dist, _ = O.match_bf(q, t)
    assert [x.trainIdx for x in m] == list(idx) and [int(x.distance) for x in m] == list(dist)
