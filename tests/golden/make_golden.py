#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the real OpenCV (cv2 4.13.0) in this container.

The reference's extractor is cv::ORB (src/frame.cc:75-79) and its frame-to-frame
matcher is cv::BFMatcher (src/pnpmatch.cc:266,278); OpenCV is an un-vendored
dependency, so the fixtures are produced by importing cv2 here.  cv2 cannot be
assumed on the GPU box, hence the committed vectors.  Extraction runs with
cv2.setUseOptimized(False): OpenCV's portable scalar path (see oracle/svo_oracle.c).

  python tests/golden/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "stereo-semantic-vo_b200"))
import cv2  # noqa: E402
import synth  # noqa: E402

KP = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"), ("octave", "<i4")])


def cv_orb(img, nf):
    cv2.setUseOptimized(False)
    kp, desc = cv2.ORB_create(nfeatures=nf, scaleFactor=1.2, nlevels=8).detectAndCompute(img, None)
    k = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=KP)
    return k, desc


def main():
    cases = []
    L, R, _ = synth.stereo_pair(synth.K_SHAPE, seed=1)
    cases.append(("k2000_seed1_left", L, 2000))
    cases.append(("k2000_seed1_right", R, 2000))
    small = synth.texture((240, 400), seed=21)
    cases.append(("s500_seed21", small, 500))
    for name, img, nf in cases:
        k, d = cv_orb(img, nf)
        # stage goldens from public cv2 calls
        lvl1 = cv2.resize(img, (int(round(img.shape[1] / 1.2)), int(round(img.shape[0] / 1.2))),
                          interpolation=cv2.INTER_LINEAR_EXACT)
        fast = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(img)
        fast = np.array([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in fast], np.int32)
        g = cv2.getGaussianKernel(7, 2, cv2.CV_32F)
        blur = cv2.sepFilter2D(img, -1, g, g, borderType=cv2.BORDER_REFLECT_101)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img, nfeatures=nf, kp=k, desc=d,
                            level1=lvl1, fast=fast, blur=blur,
                            image_sha256=hashlib.sha256(img.tobytes()).hexdigest())
        print(name, img.shape, nf, len(k))
    # BFMatcher golden with planted duplicates / ties
    rng = np.random.default_rng(5)
    q = rng.integers(0, 256, (300, 32), dtype=np.uint8)
    t = rng.integers(0, 256, (400, 32), dtype=np.uint8)
    t[100] = q[7]; t[250] = q[7]; t[30] = t[31]
    q[20] = t[30]
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
    np.savez_compressed(os.path.join(HERE, "bfmatch.npz"), q=q, t=t,
                        train=np.array([x.trainIdx for x in m], np.int32),
                        dist=np.array([x.distance for x in m], np.float32))
    print("bfmatch", len(m))


if __name__ == "__main__":
    main()
