"""Records what the REFERENCE'S OWN poseEstimationPnP (src/pnpmatch.cc:33-251, compiled unmodified into
oracle/_ref/libsvo_ref.so) decides on the hand-made adversarial feature sets of tests/adversarial_sets.py, as a golden
fixture for the GPU box (no /root/reference there): tests/test_ref_pin_adversarial.py compares both the oracle and
svo_match_greedy (CUDA) with it.  Inputs are not stored twice: the fixture holds the descriptors / positions the
reference saw (they are its outputs too) and the decisions it took.

    python tests/golden/make_golden_ref_adversarial.py        # writes tests/golden/ref_adversarial.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200"), os.path.join(ROOT, "tests")]
import adversarial_sets as A  # noqa: E402
import synth  # noqa: E402

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
BF = np.float32(CAL["bf"])


def main():
    out = {}
    for seed, boxes in A.CASES.items():
        run = A.run_reference(seed, boxes, K, BF)
        out.update(A.pack_run(run, "c%d." % seed))
        print("seed %d: %d map points created, %d claims, %d bad" % (
            seed, run["created"], int((run["cur"]["mp_create_id"] >= 0).sum()), int(run["last"]["mp_bad"].sum())))
    path = os.path.join(HERE, "ref_adversarial.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
