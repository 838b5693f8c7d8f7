"""Pins the pose-stage oracle to the REFERENCE'S OWN CODE: oracle/_ref/libsvo_ref_g2o.so is src/Optimizer.cc,
src/convert.cc and the vendored g2o (Thirdparty/g2o/g2o: core, types, stuff) compiled unmodified (oracle/Makefile
`ref_g2o`) against a stand-in for the Eigen headers this image lacks (oracle/ref_stubs_g2o/minieigen.hpp).
Optimizer::PoseOptimization (src/Optimizer.cc:15-86: one VertexSE3Expmap, EdgeSE3ProjectXYZOnlyPose edges with
Huber(sqrt(5.991)), Levenberg-Marquardt, optimize(10)) is run on a frame built through the reference's own constructors
and compared with oracle/svo_pose_oracle.c:svo_o_pose_optimize — the float32 pose the reference stores with SetPose must
be the oracle's, bit for bit."""
import numpy as np
import pytest

import synth
from oracle import oracle as O
from oracle import ref_g2o as RG

pytestmark = pytest.mark.skipif(not RG.available(), reason="/root/reference (or a prebuilt oracle/_ref/libsvo_ref_g2o.so) is not present")


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.mark.parametrize("n,seed,of,noise", [(500, 0, 0.3, 0.7), (500, 1, 0.1, 0.7), (500, 2, 0.5, 0.7), (500, 11, 0.0, 0.0),
                                             (500, 5, 0.2, 1.5), (120, 7, 0.4, 2.0), (40, 3, 0.0, 0.3), (6, 9, 0.0, 0.1)])
def test_pose_optimization_is_the_references(n, seed, of, noise):
    Xw, obs, K4, R, t, _ = synth.pose_problem(n, seed, outlier_frac=of, noise=noise)
    T0 = np.eye(4, dtype=np.float32)
    To, its, chi = O.pose_optimize(Xw, obs, K4, T0)
    Tr, ncorr = RG.pose_optimize(Xw, obs, K4, T0)
    assert ncorr == n
    assert (bits(To) == bits(Tr)).all(), np.abs(To.astype(np.float64) - Tr).max()
    if of == 0.0 and noise == 0.0:
        assert np.abs(Tr[:3, :3] - R).max() < 2e-5 and np.abs(Tr[:3, 3] - t).max() < 2e-4


def test_from_a_perturbed_pose_and_with_unmatched_keypoints():
    """A non-identity start (what Tracklastframe hands over after solvePnPRansac) and keypoints without map points
    (MapPoints[i] == NULL: src/Optimizer.cc:44 skips them)."""
    rng = np.random.default_rng(4)
    Xw, obs, K4, R, t, _ = synth.pose_problem(400, 21, outlier_frac=0.25, noise=0.8)
    w = rng.normal(0, 0.02, 3); th = np.linalg.norm(w); k = w / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    dR = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    T0 = np.eye(4, dtype=np.float32); T0[:3, :3] = (dR @ R).astype(np.float32); T0[:3, 3] = (t + rng.normal(0, 0.1, 3)).astype(np.float32)
    has = (rng.random(400) < 0.7).astype(np.uint8)
    To, its, chi = O.pose_optimize(Xw[has > 0], obs[has > 0], K4, T0)
    Tr, ncorr = RG.pose_optimize(Xw, obs, K4, T0, has=has)
    assert ncorr == int(has.sum())
    assert (bits(To) == bits(Tr)).all(), np.abs(To.astype(np.float64) - Tr).max()
