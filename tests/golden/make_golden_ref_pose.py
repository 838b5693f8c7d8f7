#!/usr/bin/env python3
"""Records what the REFERENCE'S OWN pose optimisation computes (oracle/_ref/libsvo_ref_g2o.so: src/Optimizer.cc + the vendored
g2o compiled unmodified, oracle/ref_g2o.py) on seeded problems -> tests/golden/ref_pose.npz.  /root/reference does not exist
on the GPU box, so these vectors carry the reference's answers there (tests/test_gpu_pose.py, tests/test_oracle_golden_pose.py).
usage: python tests/golden/make_golden_ref_pose.py        (needs /root/reference)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200")]
import synth                      # noqa: E402
from oracle import ref_g2o as RG  # noqa: E402

SPECS = [(500, 100, 0.3, 0.7), (500, 101, 0.0, 0.0), (300, 102, 0.5, 1.0), (60, 103, 0.1, 0.3), (12, 104, 0.0, 0.2), (450, 105, 0.2, 1.5)]
out = {"n_problems": len(SPECS)}
rng = np.random.default_rng(2024)
for i, (n, seed, of, noise) in enumerate(SPECS):
    Xw, obs, K4, R, t, _ = synth.pose_problem(n, seed, outlier_frac=of, noise=noise)
    T0 = np.eye(4, dtype=np.float32)
    if i % 2:      # a perturbed start, as after solvePnPRansac
        T0[:3, :3] = (R @ synth.rodrigues(rng.normal(0, 0.01, 3))).astype(np.float32); T0[:3, 3] = (t + rng.normal(0, 0.05, 3)).astype(np.float32)
    Tr, ncorr = RG.pose_optimize(Xw, obs, K4, T0)
    assert ncorr == n
    out["Xw%d" % i] = np.asarray(Xw, np.float32); out["obs%d" % i] = np.asarray(obs, np.float32); out["K%d" % i] = np.asarray(K4, np.float32)
    out["T0_%d" % i] = T0; out["Tref%d" % i] = Tr
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_pose.npz"), **out)
print("wrote ref_pose.npz:", len(SPECS), "problems")
