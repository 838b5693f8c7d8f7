"""Cross-checks of the sparse stereo oracle (oracle/svo_oracle.c:svo_o_stereo_sparse — north_star's ComputeStereoMatches
stage).  The reference has no sparse stereo, so the stage is DEFINED by the oracle and its parity is
unpinned; what can be checked independently is that the disparities it returns are right:

* against the ground truth of the synthetic pair (the right image is the left one warped by a known disparity field), and
* against the reference's own (uncalled) dense path frame::ElasMatch (src/frame.cc:93-120): cv::StereoSGBM with the
  reference's literal parameters, run here through cv2, read at the keypoint pixels as computekeypoint_r would
  (src/frame.cc:122-138), and
* against the dense path the reference DOES call, frame::MB (src/frame.cc:82-91): Thirdparty/MB/MSA.cpp + ctmf.c compiled
  unmodified (oracle/Makefile `ref_mb`, oracle/ref_mb_harness.cc).
"""
import ctypes
import os

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


REF_MB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libsvo_ref_mb.so")


@pytest.mark.skipif(not os.path.exists(REF_MB), reason="oracle/_ref/libsvo_ref_mb.so not built (make -C oracle ref_mb)")
@pytest.mark.parametrize("seed", [3, 8])
def test_sparse_stereo_against_the_references_own_dense_stereo(seed):
    """The stereo the reference actually runs (src/Tracking.cc:226-228): frame::MB (src/frame.cc:82-91) ->
    Thirdparty/MB/MSA.cpp + ctmf.c, compiled unmodified into oracle/_ref/libsvo_ref_mb.so, 48 integer disparity levels.
    Read at the keypoints as computekeypoint_r does (src/frame.cc:131: dispimg.at<float>(ly, lx), i.e. truncated
    coordinates), it agrees with the sparse oracle's sub-pixel disparity to within its own integer quantisation on most
    keypoints.  This is a cross-check of the disparities, not a parity pin: the two algorithms differ by design."""
    bf = CAL["bf"]; b = bf / CAL["fx"]
    L, R, _ = synth.stereo_pair(seed=seed)
    h, w = L.shape
    lib = ctypes.CDLL(REF_MB)
    bgr = [np.ascontiguousarray(np.repeat(im[:, :, None], 3, 2)) for im in (L, R)]     # MSA::init reads 3 bytes / pixel
    D = np.full((h, w), -1, np.float32)
    vp = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    assert lib.ref_mb_disparity(vp(bgr[0]), vp(bgr[1]), w, h, vp(D)) == 0
    assert D.min() >= 0 and D.max() <= 48 and (D == np.rint(D)).all()                  # integer levels of MSA::solve
    kl, dl, pl = O.orb(L, 2000, with_pyramid=True)
    kr, dr, pr = O.orb(R, 2000, with_pyramid=True)
    ur, dep, mr, sad = O.stereo_sparse(kl, dl, pl, kr, dr, pr, bf, b)
    O.pyramid_free(pl); O.pyramid_free(pr)
    ok = dep > 0
    x, y = kl["x"], kl["y"]
    ds = D[y.astype(int), x.astype(int)]
    both = ok & (ds > 0)
    assert both.sum() >= 600
    e = np.abs((x - ur) - ds)[both]
    assert np.median(e) <= 0.5 and (e < 1.0).mean() > 0.75 and (e < 2.0).mean() > 0.82, \
        (np.median(e), (e < 1.0).mean(), (e < 2.0).mean())
