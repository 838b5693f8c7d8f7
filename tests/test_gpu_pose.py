"""GPU parity of the pose stage (csrc/pose.cu through the C ABI) against oracle/svo_pose_oracle.c.

Tolerances (floating point; north_star names none for this stage): the LM and the RANSAC refit accumulate
their 6x6 systems as float64 block reductions instead of the oracle's sequential sums, so poses agree to
~1e-12; the tests allow 1e-6 on R/t (1e-5 on the float32 Tcw the LM returns).  Everything discrete —
winning sample, its solution index, inlier flags — must be equal.  The LM's outer-iteration count is equal
too except at convergence, where g2o's accept test compares two costs that differ by rounding noise: there
the count may differ while pose and cost still agree.
"""
import os

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ctx():
    import svo as S
    c = S.Context(1241, 376, nfeatures=2000, max_batch=8, lanes=1, max_rows=5000)
    yield c
    c.close()


def problems(specs):
    out = []
    for n, seed, of, noise in specs:
        Xw, obs, K4, R, t, bad = synth.pose_problem(n, seed, of, noise)
        out.append(dict(pts3d=Xw, pts2d=obs, K=K4, R=R, t=t))
    return out


SPECS = [(800, 0, 0.3, 0.5), (2000, 1, 0.2, 0.5), (300, 2, 0.5, 1.0), (60, 3, 0.1, 0.3), (1500, 4, 0.0, 0.0),
         (5, 5, 0.0, 0.1), (3, 6, 0.0, 0.0), (0, 7, 0.0, 0.0)]


def test_pnp_ransac_matches_oracle_batch(ctx):
    pr = problems(SPECS)
    for seed in (1, 77):
        res = ctx.pnp_ransac(pr, iterations=100, reproj_err=8.0, seed=seed, refine_iters=10)
        for q, r in zip(pr, res):
            n, R, t, mask, info = O.pnp_ransac(q["pts3d"], q["pts2d"], q["K"], 100, 8.0, seed, 10)
            assert r["n_inliers"] == n, (len(q["pts3d"]), r["n_inliers"], n)
            assert tuple(r["info"]) == tuple(int(v) for v in info)
            assert (r["inliers"] == mask).all()
            if n:
                assert np.abs(r["R"] - R).max() < 1e-6 and np.abs(r["t"] - t).max() < 1e-6
                if len(q["pts3d"]) >= 60:
                    assert np.abs(r["R"] - q["R"]).max() < 3e-3


def test_pnp_ransac_more_iterations_and_thresholds(ctx):
    pr = problems([(1200, 31, 0.6, 0.8), (400, 32, 0.7, 0.5)])
    for iters, thr in ((512, 2.0), (37, 8.0), (1, 8.0)):
        res = ctx.pnp_ransac(pr, iterations=iters, reproj_err=thr, seed=5, refine_iters=4)
        for q, r in zip(pr, res):
            n, R, t, mask, info = O.pnp_ransac(q["pts3d"], q["pts2d"], q["K"], iters, thr, 5, 4)
            assert r["n_inliers"] == n and tuple(r["info"]) == tuple(int(v) for v in info)
            assert (r["inliers"] == mask).all()
            if n:
                assert np.abs(r["R"] - R).max() < 1e-6 and np.abs(r["t"] - t).max() < 1e-6


def test_pnp_ransac_against_opencv_golden(ctx):
    z = np.load(os.path.join(G, "pose.npz"))
    pr = [dict(pts3d=z["Xw%d" % i], pts2d=z["obs%d" % i], K=tuple(z["K%d" % i])) for i in range(int(z["ncases"]))]
    res = ctx.pnp_ransac(pr)
    for i, r in enumerate(res):
        cvm = z["cvmask%d" % i].astype(bool); m = r["inliers"].astype(bool)
        assert (m & cvm).sum() / max(1, (m | cvm).sum()) >= 0.98
        assert np.abs(r["R"] - z["cvR%d" % i]).max() < 1e-4 and np.abs(r["t"] - z["cvt%d" % i]).max() < 2e-3


def test_pnp_ransac_points_beyond_shared_memory_and_device_pointers(ctx):
    # 9000 points do not fit the shared-memory staging: the kernel reads them from global memory
    Xw, obs, K4, R, t, _ = synth.pose_problem(9000, 40, 0.25, 0.5)
    d3, d2 = ctx.to_device(Xw), ctx.to_device(obs)
    import svo as S
    arr = (S.PoseProblem * 1)()
    arr[0].pts3d, arr[0].pts2d, arr[0].n = d3, d2, len(Xw)
    arr[0].fx, arr[0].fy, arr[0].cx, arr[0].cy = [float(v) for v in K4]
    res = (S.PnpResult * 1)()
    mask = np.zeros(len(Xw), np.uint8)
    ctx._chk(ctx.lib.svo_pnp_ransac(ctx.h, arr, 1, 100, 8.0, 3, 10, res, mask.ctypes.data))
    n, Ro, to, mo, info = O.pnp_ransac(Xw, obs, K4, 100, 8.0, 3, 10)
    assert res[0].n_inliers == n and (mask == mo).all()
    assert np.abs(np.array(res[0].R[:]).reshape(3, 3) - Ro).max() < 1e-6
    ctx.lib.svo_free_device(ctx.h, d3); ctx.lib.svo_free_device(ctx.h, d2)


def test_pose_optimize_matches_oracle_batch(ctx):
    pr = problems(SPECS)
    rng = np.random.default_rng(0)
    for q in pr:
        T0 = np.eye(4, dtype=np.float32)
        T0[:3, :3] = q["R"] @ synth.rodrigues(rng.normal(0, 0.01, 3)); T0[:3, 3] = q["t"] + rng.normal(0, 0.05, 3)
        q["Tcw"] = T0
    for iters in (10, 1, 0, 3):
        res = ctx.pose_optimize(pr, iterations=iters)
        for q, (T, its, chi) in zip(pr, res):
            To, io, co = O.pose_optimize(q["pts3d"], q["pts2d"], q["K"], q["Tcw"], iters)
            assert np.abs(T - To).max() < 1e-5
            assert abs(chi - co) <= 1e-9 * max(1.0, abs(co))
            assert its == io or iters == 10, (len(q["pts3d"]), iters, its, io)


def test_pose_optimize_against_the_references_recorded_poses(ctx):
    """tests/golden/ref_pose.npz holds what the reference's OWN Optimizer::PoseOptimization (src/Optimizer.cc + the vendored
    g2o, compiled unmodified: oracle/_ref/libsvo_ref_g2o.so, tests/golden/make_golden_ref_pose.py) stored with SetPose;
    the oracle reproduces those float32 poses bit for bit (tests/test_ref_pin_pose.py), the kernel to the tolerance above."""
    z = np.load(os.path.join(G, "ref_pose.npz"))
    pr = [dict(pts3d=z["Xw%d" % i], pts2d=z["obs%d" % i], K=z["K%d" % i], Tcw=z["T0_%d" % i]) for i in range(int(z["n_problems"]))]
    res = ctx.pose_optimize(pr)
    for i, (T, its, chi) in enumerate(res):
        assert np.abs(T - z["Tref%d" % i]).max() < 1e-5, i


def test_pose_optimize_from_identity_and_after_ransac(ctx):
    pr = problems([(1000, 50, 0.3, 0.5), (1000, 51, 0.0, 0.0)])
    for q in pr:
        q["Tcw"] = np.eye(4, dtype=np.float32)
    res = ctx.pose_optimize(pr)
    for q, (T, its, chi) in zip(pr, res):
        To, io, co = O.pose_optimize(q["pts3d"], q["pts2d"], q["K"], q["Tcw"])
        assert np.abs(T - To).max() < 1e-5 and abs(chi - co) <= 1e-9 * max(1.0, abs(co)) and abs(its - io) <= 2
    assert np.abs(res[1][0][:3, :3] - pr[1]["R"]).max() < 2e-5     # exact data -> exact pose
    # the reference's order: solvePnPRansac, SetPose, then PoseOptimization over all matches (src/Tracking.cc:114-120)
    rr = ctx.pnp_ransac(pr)
    for q, r in zip(pr, rr):
        T = np.eye(4, dtype=np.float32); T[:3, :3] = r["R"]; T[:3, 3] = r["t"]
        q["Tcw"] = T
    res = ctx.pose_optimize(pr)
    for q, (T, its, chi) in zip(pr, res):
        To, io, co = O.pose_optimize(q["pts3d"], q["pts2d"], q["K"], q["Tcw"])
        assert np.abs(T - To).max() < 1e-5 and abs(chi - co) <= 1e-9 * max(1.0, abs(co)) and abs(its - io) <= 2


def test_pose_capacity_and_argument_errors(ctx):
    import svo as S
    pr = problems([(10, 0, 0, 0)] * 9)
    with pytest.raises(S.SvoError) as e:
        ctx.pnp_ransac(pr)
    assert e.value.code == S.E_CAPACITY
    with pytest.raises(S.SvoError) as e:
        ctx.pnp_ransac(pr[:1], iterations=100000)
    assert e.value.code == S.E_CAPACITY
    with pytest.raises(S.SvoError) as e:
        ctx.pnp_ransac(pr[:1], reproj_err=-1.0)
    assert e.value.code == S.E_INVALID
