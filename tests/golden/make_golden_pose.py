#!/usr/bin/env python3
"""Generate tests/golden/pose.npz: inputs of the pose stage with what OpenCV 4.13's
cv2.solvePnPRansac (the runnable stand-in for the call at src/pnpmatch.cc:227) returns on them.

The GPU stage is NOT a transcription of OpenCV's RANSAC (see oracle/svo_pose_oracle.c), so these
vectors pin agreement statistically: pose within tolerance, inlier sets overlapping.

  python tests/golden/make_golden_pose.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "stereo-semantic-vo_b200"))
import cv2  # noqa: E402
import synth  # noqa: E402

CASES = [(800, 0, 0.3, 0.5), (2000, 1, 0.2, 0.5), (300, 2, 0.5, 1.0), (60, 3, 0.1, 0.3), (1500, 4, 0.0, 0.0)]


def main():
    out = {}
    for i, (n, seed, of, noise) in enumerate(CASES):
        Xw, obs, K4, R, t, bad = synth.pose_problem(n, seed, of, noise)
        Kc = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
        cv2.setRNGSeed(0)
        ok, rv, tv, inl = cv2.solvePnPRansac(Xw.astype(np.float64), obs.astype(np.float64), Kc, None,
                                             iterationsCount=100, reprojectionError=8.0, confidence=0.99)
        assert ok
        Rc, _ = cv2.Rodrigues(rv)
        mask = np.zeros(n, np.uint8); mask[inl.ravel()] = 1
        out.update({"Xw%d" % i: Xw, "obs%d" % i: obs, "K%d" % i: np.array(K4, np.float64), "Rtrue%d" % i: R, "ttrue%d" % i: t,
                    "cvR%d" % i: Rc, "cvt%d" % i: tv.ravel(), "cvmask%d" % i: mask})
        print(i, n, "inliers", int(mask.sum()), "of", int((~bad).sum()), "clean")
    np.savez_compressed(os.path.join(HERE, "pose.npz"), ncases=len(CASES), **out)


if __name__ == "__main__":
    main()
