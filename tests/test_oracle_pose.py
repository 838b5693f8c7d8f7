"""CPU checks of the pose-stage oracle (oracle/svo_pose_oracle.c).

* svo_o_pose_optimize restates Optimizer::PoseOptimization (src/Optimizer.cc:15-86) from the vendored g2o
  sources.  It is pinned bit for bit to the reference's own build of that function in tests/test_ref_pin_pose.py
  (and against recorded vectors in tests/test_oracle_golden_pose.py); here it is checked through properties:
  exact data gives the exact pose, the robust cost never increases and ends at a stationary point of the
  Huber cost (gradient computed independently in numpy), zero edges leave the pose alone.
* svo_o_pnp_ransac DEFINES the data-parallel stand-in for cv::solvePnPRansac (src/pnpmatch.cc:227); it is
  compared with OpenCV 4.13's answer on committed vectors (tests/golden/pose.npz, made by
  tests/golden/make_golden_pose.py) and with cv2 live when it is importable.
"""
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
DELTA = float(np.float32(np.sqrt(5.991)))


def T_of(R, t):
    T = np.eye(4, dtype=np.float32); T[:3, :3] = R; T[:3, 3] = t
    return T


def huber_cost_and_grad(Xw, obs, K4, T):
    """robust chi2 and its gradient w.r.t. a left-multiplied se(3) increment (omega, upsilon) — numpy, independent."""
    fx, fy, cx, cy = K4
    R = T[:3, :3].astype(np.float64); t = T[:3, 3].astype(np.float64)
    Xc = Xw.astype(np.float64) @ R.T + t
    x, y, z = Xc.T
    e = np.stack([obs[:, 0] - (fx * x / z + cx), obs[:, 1] - (fy * y / z + cy)], 1)
    chi = (e ** 2).sum(1)
    w = np.where(chi <= DELTA ** 2, 1.0, DELTA / np.sqrt(np.maximum(chi, 1e-300)))
    rho = np.where(chi <= DELTA ** 2, chi, 2 * np.sqrt(chi) * DELTA - DELTA ** 2)
    # d proj / d Xc, and d Xc / d(omega, upsilon) = [-[Xc]x, I]
    g = np.zeros(6)
    for i in range(len(Xw)):
        Jp = np.array([[fx / z[i], 0, -fx * x[i] / z[i] ** 2], [0, fy / z[i], -fy * y[i] / z[i] ** 2]])
        Xx = np.array([[0, -z[i], y[i]], [z[i], 0, -x[i]], [-y[i], x[i], 0]])
        J = -Jp @ np.hstack([-Xx, np.eye(3)])          # d e / d increment
        g += 2 * w[i] * (J.T @ e[i])
    return rho.sum(), g


def test_lm_exact_data_recovers_pose():
    Xw, obs, K4, R, t, _ = synth.pose_problem(500, 11, outlier_frac=0.0, noise=0.0)
    T, its, chi = O.pose_optimize(Xw, obs, K4, np.eye(4, dtype=np.float32))
    assert np.abs(T[:3, :3] - R).max() < 2e-5 and np.abs(T[:3, 3] - t).max() < 2e-4
    assert chi < 1e-3 * len(Xw)       # float32 observations: ~1e-5 px residuals
    assert 1 <= its <= 10


@pytest.mark.parametrize("seed,of", [(0, 0.3), (1, 0.1), (2, 0.5)])
def test_lm_descends_to_a_stationary_point_of_the_huber_cost(seed, of):
    Xw, obs, K4, R, t, _ = synth.pose_problem(600, seed, outlier_frac=of, noise=0.7)
    T0 = np.eye(4, dtype=np.float32)
    c0, g0 = huber_cost_and_grad(Xw, obs, K4, T0)
    prev = c0
    for iters in (1, 2, 3, 5, 10):
        T, its, chi = O.pose_optimize(Xw, obs, K4, T0, iterations=iters)
        c, _ = huber_cost_and_grad(Xw, obs, K4, T)
        assert abs(c - chi) <= 1e-3 * max(1.0, chi)     # the pose is rounded to float32 on the way out
        assert c <= prev * (1 + 1e-9)
        prev = c
    c, g = huber_cost_and_grad(Xw, obs, K4, T)
    # g2o stops early once the relative gain stays under 1e-3 three times ("Stop criterium (Raul)"), and the pose
    # leaves as float32: the gradient shrinks by orders of magnitude, not to zero
    assert np.linalg.norm(g) < 5e-2 * np.linalg.norm(g0)


def test_lm_zero_edges_and_iterations_keep_the_pose():
    R = synth.rodrigues([0.1, -0.2, 0.05]); T0 = T_of(R, [1, 2, 3])
    T, its, chi = O.pose_optimize(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), (700, 700, 600, 180), T0)
    assert np.abs(T - T0).max() < 1e-6 and chi == 0
    Xw, obs, K4, *_ = synth.pose_problem(50, 3)
    T, its, chi = O.pose_optimize(Xw, obs, K4, T0, iterations=0)
    assert np.abs(T - T0).max() < 1e-6 and its == 0


def test_ransac_matches_opencv_golden():
    z = np.load(os.path.join(G, "pose.npz"))
    for i in range(int(z["ncases"])):
        Xw, obs, K4 = z["Xw%d" % i], z["obs%d" % i], tuple(z["K%d" % i])
        n, R, t, mask, info = O.pnp_ransac(Xw, obs, K4, iterations=100, reproj_err=8.0, seed=1)
        cvm = z["cvmask%d" % i].astype(bool); m = mask.astype(bool)
        jac = (m & cvm).sum() / max(1, (m | cvm).sum())
        assert jac >= 0.98, (i, jac)
        assert n == m.sum()
        # both refits minimise the squared reprojection error over (nearly) the same inliers
        assert np.abs(R - z["cvR%d" % i]).max() < 1e-4, i
        assert np.abs(t - z["cvt%d" % i]).max() < 2e-3, i
        assert np.abs(R - z["Rtrue%d" % i]).max() < 2e-3 and np.abs(t - z["ttrue%d" % i]).max() < 2e-2


def test_ransac_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    for seed in range(20, 26):
        Xw, obs, K4, R, t, bad = synth.pose_problem(700, seed, outlier_frac=0.35, noise=0.5)
        Kc = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float64)
        ok, rv, tv, inl = cv2.solvePnPRansac(Xw.astype(np.float64), obs.astype(np.float64), Kc, None,
                                             iterationsCount=100, reprojectionError=8.0, confidence=0.99)
        n, Ro, to, mask, _ = O.pnp_ransac(Xw, obs, K4, seed=seed)
        cvm = np.zeros(len(Xw), bool); cvm[inl.ravel()] = True
        m = mask.astype(bool)
        assert (m & cvm).sum() / (m | cvm).sum() >= 0.98
        assert np.abs(Ro - cv2.Rodrigues(rv)[0]).max() < 1e-4 and np.abs(to - tv.ravel()).max() < 2e-3


def test_ransac_exact_data_and_seed_independence():
    Xw, obs, K4, R, t, _ = synth.pose_problem(400, 7, outlier_frac=0.0, noise=0.0)
    for seed in (1, 2, 99):
        n, Ro, to, mask, info = O.pnp_ransac(Xw, obs, K4, seed=seed)
        assert n == 400 and mask.all()
        assert np.abs(Ro - R).max() < 1e-5 and np.abs(to - t).max() < 1e-4


def test_ransac_degenerate_inputs():
    K4 = (700.0, 700.0, 600.0, 180.0)
    n, R, t, mask, info = O.pnp_ransac(np.zeros((0, 3), np.float32), np.zeros((0, 2), np.float32), K4)
    assert n == 0
    n, R, t, mask, info = O.pnp_ransac(np.ones((2, 3), np.float32), np.ones((2, 2), np.float32), K4)
    assert n == 0 and not mask.any()
    # collinear world points: no triangle, no model
    X = np.stack([np.linspace(-1, 1, 30), np.zeros(30), np.full(30, 10.0)], 1).astype(np.float32)
    o = np.stack([K4[0] * X[:, 0] / X[:, 2] + K4[2], np.full(30, K4[3])], 1).astype(np.float32)
    n, R, t, mask, info = O.pnp_ransac(X, o, K4)
    assert n == 0 and info[2] == 0
    # pure outliers: whatever wins has few inliers and the call stays finite
    rng = np.random.default_rng(0)
    X = rng.uniform(-5, 5, (200, 3)).astype(np.float32) + np.float32([0, 0, 20])
    o = np.stack([rng.uniform(0, 1241, 200), rng.uniform(0, 376, 200)], 1).astype(np.float32)
    n, R, t, mask, info = O.pnp_ransac(X, o, K4)
    assert n < 40 and np.isfinite(R).all() and np.isfinite(t).all()
