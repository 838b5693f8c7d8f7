"""The C++ drop-in (frame / pnpmatch mirror, stereo-semantic-vo_b200/adapter) driven the way
Tracking::Track drives the reference (src/Tracking.cc:225-238), checked against the oracle."""
import os
import subprocess

import numpy as np
import pytest

import synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADAPTER = os.path.join(ROOT, "stereo-semantic-vo_b200", "adapter")


def lookup_depth(kp, ur, dep, bf, H, W):
    """What the tracker reads: disparity scattered at keypoint pixels in keypoint order, then
    depthimg = bf / dispimg looked up at each keypoint's pixel (src/frame.cc:122-164, Tracking.cc:51)."""
    disp = np.full((H, W), -1, np.float32)
    for i in np.nonzero(dep > 0)[0]:
        disp[int(kp["y"][i]), int(kp["x"][i])] = kp["x"][i] - ur[i]
    d = disp[kp["y"].astype(np.int64), kp["x"].astype(np.int64)]
    with np.errstate(divide="ignore"):
        return np.where(d != 0, np.float32(bf) / d, np.float32(-1)).astype(np.float32)


def expected_keypoints_r(kp, ur, dep, H, W):
    """frame::computekeypoint_r over the scattered disparity image (rx sticks when disp == -1)."""
    disp = np.full((H, W), -1, np.float32)
    for i in np.nonzero(dep > 0)[0]:
        disp[int(kp["y"][i]), int(kp["x"][i])] = kp["x"][i] - ur[i]
    out = np.empty(len(kp), np.float32)
    rx = np.float32(-1)
    for i in range(len(kp)):
        d = disp[int(kp["y"][i]), int(kp["x"][i])]
        if d != -1:
            rx = kp["x"][i] - d
        out[i] = rx
    return out


def test_adapter_two_frames(tmp_path):
    exe = os.path.join(ADAPTER, "adapter_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    H, W, NF = 376, 1241, 1000
    seq = synth.Sequence((H, W), seed=5)
    (L0, R0), (L1, R1) = seq.frame(0), seq.frame(1)
    paths = []
    for name, im in (("L0", L0), ("R0", R0), ("L1", L1), ("R1", R1)):
        p = str(tmp_path / (name + ".raw"))
        im.tofile(p)
        paths.append(p)
    out = str(tmp_path / "out.txt")
    subprocess.check_call([exe, str(W), str(H), str(NF)] + paths + [out])
    kp_rows, map_rank, scores, pairs, pnp, lm = [], {}, {}, [], None, None
    for line in open(out):
        t = line.split()
        if t[0] == "counts":
            counts = list(map(int, t[1:]))
        elif t[0] == "kp":
            kp_rows.append(t[1:])
        elif t[0] == "map":
            map_rank[int(t[1])] = int(t[2])
        elif t[0] == "score":
            scores[int(t[1])] = np.float32(t[2])
        elif t[0] == "pair":
            pairs.append([np.float32(v) for v in t[2:7]])
        elif t[0] == "pnp":
            pnp = t[1:]
        elif t[0] == "lm":
            lm = t[1:]
    cal = synth.KITTI_04_12
    bf = np.float32(379.8145)
    b = np.float32(bf / np.float32(707.0912))
    k0, d0, p0 = O.orb(L0, NF, with_pyramid=True)
    k0r, d0r, p0r = O.orb(R0, NF, with_pyramid=True)
    ur0, dep0, _, _ = O.stereo_sparse(k0, d0, p0, k0r, d0r, p0r, bf, b)
    k1, d1, p1 = O.orb(L1, NF, with_pyramid=True)
    k1r, d1r, p1r = O.orb(R1, NF, with_pyramid=True)
    ur1, dep1, _, _ = O.stereo_sparse(k1, d1, p1, k1r, d1r, p1r, bf, b)
    assert counts[0] == len(k0) and counts[1] == len(k1) == len(kp_rows)
    got = np.array([[np.float32(v) for v in r[1:5]] for r in kp_rows], np.float32)
    assert (got[:, 0].view(np.uint32) == k1["x"].view(np.uint32)).all()
    assert (got[:, 1].view(np.uint32) == k1["y"].view(np.uint32)).all()
    assert (got[:, 2].view(np.uint32) == k1["angle"].view(np.uint32)).all()
    assert (got[:, 3].view(np.uint32) == k1["response"].view(np.uint32)).all()
    assert ([int(r[5]) for r in kp_rows] == k1["octave"]).all()
    desc = np.array([[int(v) for v in r[9:41]] for r in kp_rows], np.uint8)
    assert (desc == d1).all()
    # stereo fields: keypoints_r.x and depthimg at the keypoint
    z = np.array([np.float32(r[7]) for r in kp_rows]); xr = np.array([np.float32(r[6]) for r in kp_rows])
    z1 = lookup_depth(k1, ur1, dep1, bf, H, W)
    valid = z1 > 0
    assert ((z > 0) == valid).all()
    assert (np.abs(z[valid] - z1[valid]) <= 2e-3 * z1[valid]).all()
    exp_xr = expected_keypoints_r(k1, ur1, dep1, H, W)
    assert np.abs(xr - exp_xr).max() <= 2e-3
    assert (dep1 > 0).sum() > 200
    # pass 1: rows = last frame's keypoints that own a map point (depthimg > 0 at their pixel)
    live = (lookup_depth(k0, ur0, dep0, bf, H, W) > 0).astype(np.uint8)
    g1 = O.match_greedy(d0, d1, 0, row_live=live)
    assert counts[2] == int(g1["row_claimed"].sum()) and counts[2] > 50
    for i, s in scores.items():
        with np.errstate(divide="ignore"):
            assert s == np.float32(g1["second"][i]) / np.float32(g1["best"][i])
    # pass 2: the local map in the std::set's own order, minus the points already observing the frame
    order = [map_rank[r] for r in sorted(map_rank)]
    rows2 = [i for i in order if not g1["row_claimed"][i]]
    g2 = O.match_greedy(d0[rows2], d1, 1, claimed=g1["claimed"], claim_row=g1["claim_row"], row_base=10 ** 6)
    assert counts[3] == int(g2["row_claimed"].sum())
    expect = g1["claim_row"].copy()
    for j in np.nonzero(g2["claim_row"] >= 10 ** 6)[0]:
        expect[j] = rows2[g2["claim_row"][j] - 10 ** 6]
    assert ([int(r[8]) for r in kp_rows] == expect).all()
    # pose stage: device PnP RANSAC on the matched 3D-2D pairs, SetPose, Optimizer::PoseOptimization
    pairs = np.array(pairs, np.float32)
    nm = int((expect >= 0).sum())
    assert len(pairs) == nm and int(lm[0]) == nm
    assert (pairs[:, 3] == k1["x"][expect >= 0]).all() and (pairs[:, 4] == k1["y"][expect >= 0]).all()
    K4 = (np.float32(707.0912), np.float32(707.0912), np.float32(601.8873), np.float32(183.1104))
    n, Ro, to, mask, info = O.pnp_ransac(pairs[:, :3], pairs[:, 3:], K4, 100, 8.0, 1, 10)
    assert int(pnp[0]) == 1 and int(pnp[1]) == n and n > 30
    Tcl = np.array([np.float32(v) for v in pnp[2:18]], np.float32).reshape(4, 4)
    assert np.abs(Tcl[:3, :3] - Ro).max() < 1e-5 and np.abs(Tcl[:3, 3] - to).max() < 1e-5
    To, its, chi = O.pose_optimize(pairs[:, :3], pairs[:, 3:], K4, Tcl)
    Tlm = np.array([np.float32(v) for v in lm[1:17]], np.float32).reshape(4, 4)
    assert np.abs(Tlm - To).max() < 1e-5
    for p in (p0, p0r, p1, p1r):
        O.pyramid_free(p)
