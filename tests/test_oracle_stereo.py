"""Cross-checks of the sparse stereo oracle (oracle/svo_oracle.c:svo_o_stereo_sparse — north_star's ComputeStereoMatches
stage).  The reference has no sparse stereo, so the stage is DEFINED by the oracle and its parity is
unpinned; what can be checked independently is that the disparities it returns are right:

* against the ground truth of the synthetic pair (the right image is the left one warped by a known disparity field), and
* against the reference's own (uncalled) dense path frame::ElasMatch (src/frame.cc:93-120): cv::StereoSGBM with the
  reference's literal parameters, run here through cv2, read at the keypoint pixels as computekeypoint_r would
  (src/frame.cc:122-138).
"""
import cv2
import numpy as np
import pytest

import synth
from oracle import oracle as O

CAL = synth.KITTI_04_12


def elas_match(L, R):
    """frame::ElasMatch (src/frame.cc:93-120) through cv2: same constructor arguments and setters, disparity / 16."""
    nd = ((L.shape[0] // 8) + 15) & -16
    sg = cv2.StereoSGBM_create(0, 16, 3)
    sg.setPreFilterCap(63)
    win = 9
    sg.setBlockSize(win)
    sg.setP1(8 * 1 * win * win); sg.setP2(32 * 1 * win * win)
    sg.setMinDisparity(0); sg.setNumDisparities(nd)
    sg.setUniquenessRatio(10); sg.setSpeckleWindowSize(100); sg.setSpeckleRange(32); sg.setDisp12MaxDiff(1)
    sg.setMode(cv2.StereoSGBM_MODE_SGBM)
    return sg.compute(L, R).astype(np.float32) * np.float32(1.0 / 16)


@pytest.mark.parametrize("seed", [3, 8])
def test_sparse_stereo_against_ground_truth_and_the_references_sgbm(seed):
    bf = CAL["bf"]; b = bf / CAL["fx"]
    L, R, disp = synth.stereo_pair(seed=seed)
    kl, dl, pl = O.orb(L, 2000, with_pyramid=True)
    kr, dr, pr = O.orb(R, 2000, with_pyramid=True)
    ur, dep, mr, sad = O.stereo_sparse(kl, dl, pl, kr, dr, pr, bf, b)
    O.pyramid_free(pl); O.pyramid_free(pr)
    ok = dep > 0
    assert ok.sum() >= 600
    x, y = kl["x"], kl["y"]
    d = x - ur
    assert (np.abs(dep[ok] - np.float32(bf) / d[ok]) <= 1e-3 * dep[ok]).all()          # depth = bf / disparity
    yi = np.clip(np.rint(y).astype(int), 0, L.shape[0] - 1)
    # ground truth: right(xr) = left(xr + disp(xr)), i.e. the field is indexed at the RIGHT pixel
    truth = disp[yi, np.clip(np.rint(ur).astype(int), 0, L.shape[1] - 1)]
    e = np.abs(d - truth)[ok]
    assert np.median(e) < 0.4 and (e < 1.0).mean() > 0.85, (np.median(e), (e < 1.0).mean())
    # the reference's dense path at the keypoint pixels
    D = elas_match(L, R)
    ds = D[yi, np.clip(np.rint(x).astype(int), 0, L.shape[1] - 1)]
    both = ok & (ds > 0)
    assert both.sum() >= 500
    e2 = np.abs(d - ds)[both]
    assert np.median(e2) < 0.4 and (e2 < 1.0).mean() > 0.85, (np.median(e2), (e2 < 1.0).mean())
