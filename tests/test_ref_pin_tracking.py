"""Pins oracle/track.py — the oracle of the device-resident tracker state (csrc/track.cu) — to the reference's OWN
Tracking::Track: oracle/_ref/libsvo_ref_g2o.so holds src/Tracking.cc compiled unmodified together with src/frame.cc,
src/mappoint.cc, src/pnpmatch.cc, src/Optimizer.cc, src/convert.cc and the vendored g2o (Pangolin and the viewer are
stubbed; frame::MB's dense solver returns the synthetic disparity of the frame; OpenCV calls are answered by cv2).  Seven
frames go through Tracking::Track exactly as main.cpp:176 drives it — including Tracking::init on frame 0 and
Optimizer::PoseOptimization on every frame — and every frame's CurrentFrame->MapPoints, match_score, bad flags, created
points and LocalMapPoints are compared with the tracker oracle (the same checks tests/test_oracle_track.py applies to the
loop its harness restates)."""
import numpy as np
import pytest

import synth
from oracle import ref_g2o as RG
from test_oracle_track import BF, BOXES, K, check, replay

pytestmark = pytest.mark.skipif(not RG.available(), reason="/root/reference (or a prebuilt oracle/_ref/libsvo_ref_g2o.so) is not present")


def tracking_run(seed, n, boxes_of, tmpdir):
    seq = synth.Sequence(synth.K_SHAPE, seed=seed)
    frames = [seq.frame(t) for t in range(n)]
    disps = [synth.dense_disparity(synth.K_SHAPE, 100 * seed + t) for t in range(n)]
    return RG.run_tracking(frames, disps, K, BF, [boxes_of(t) for t in range(n)], tmpdir)


def test_tracker_oracle_follows_tracking_track_over_seven_frames(tmp_path):
    boxes_of = lambda t: BOXES if t % 2 == 1 else BOXES[:1]
    recs = tracking_run(5, 7, boxes_of, tmp_path)
    n_bad, n_p1, n_p2 = check(recs, replay(recs, boxes_of))
    assert n_p1 > 300 and n_p2 > 10 and n_bad >= 3
    assert sum(r["erased"] for r in recs) > 100, "the 4-frame window must have dropped points"
    # Optimizer::PoseOptimization ran on every frame: the poses moved away from the identity
    assert any(not np.allclose(r["last"]["Tcw"], np.eye(4), atol=1e-4) for r in recs[1:])


def test_tracker_oracle_follows_tracking_track_without_boxes(tmp_path):
    boxes_of = lambda t: []
    recs = tracking_run(6, 5, boxes_of, tmp_path)
    n_bad, n_p1, n_p2 = check(recs, replay(recs, boxes_of))
    assert n_bad == 0 and n_p1 > 200
