"""cv2.ORB_create(nf, 1.2, 8).detectAndCompute on images whose pyramid level sizes depend on how cvRound(cols / scale) is
evaluated (249 x 181 -> level 1 is 208 wide in cv2 4.13, the quotient would give 207; 465 x 297 -> 388 x 248), recorded
for the GPU box: tests/test_gpu_parity2.py compares svo_extract with them.  Images come from synth.texture(shape, 77).

    python tests/golden/make_golden_half_sizes.py        # writes tests/golden/half_sizes.npz
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200")]
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [((181, 249), 200), ((297, 465), 600)]


def main():
    cv2.setUseOptimized(False)
    out = {}
    for (h, w), nf in CASES:
        img = synth.texture((h, w), 77)
        kp, desc = cv2.ORB_create(nfeatures=nf, scaleFactor=1.2, nlevels=8).detectAndCompute(img, None)
        out["kp_%dx%d" % (w, h)] = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=O.KP_DTYPE)
        out["desc_%dx%d" % (w, h)] = desc
        print(w, h, nf, len(kp), "keypoints")
    path = os.path.join(HERE, "half_sizes.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
