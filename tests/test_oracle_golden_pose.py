"""The pose-stage oracle against vectors recorded from the REFERENCE'S OWN code (tests/golden/ref_pose.npz, written by
tests/golden/make_golden_ref_pose.py from oracle/_ref/libsvo_ref_g2o.so = src/Optimizer.cc + the vendored g2o compiled
unmodified): svo_o_pose_optimize returns the float32 pose Optimizer::PoseOptimization stored, bit for bit.  Runs without
/root/reference (tests/test_ref_pin_pose.py is the live comparison)."""
import os

import numpy as np

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_equals_the_recorded_reference_poses():
    z = np.load(os.path.join(G, "ref_pose.npz"))
    for i in range(int(z["n_problems"])):
        To, its, chi = O.pose_optimize(z["Xw%d" % i], z["obs%d" % i], z["K%d" % i], z["T0_%d" % i])
        assert (To.view(np.uint32) == z["Tref%d" % i].view(np.uint32)).all(), i
