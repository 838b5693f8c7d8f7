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


# ---- the drop-in entry points exactly as Tracking::Tracklastframe calls them, against the reference's own code ----
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.mark.parametrize("seed", [5, 9])
def test_adapter_pose_estimation_pnp_vs_reference_golden(tmp_path, seed):
    """pnpmatch::poseEstimationPnP / poseEstimation2D_2D / find_feature_matches through the adapter, with non-empty
    offline_box (src/Tracking.cc:114, src/pnpmatch.cc:33-251, :253-337), against tests/golden/ref_track_seed*.npz —
    what the reference's own compiled code (oracle/_ref) produced on the same inputs: ORB output after the re-extraction,
    keypoints_r, depthimg at the keypoints, createmappoint's selection and world positions (UnprojectStereo), the
    point lists handed to findFundamentalMat, match_score and the map points pass 1 turned bad.  The final
    CurrentFrame->MapPoints depends on the std::set's pointer order in pass 2, so it is checked against the oracle
    (pinned to the reference in tests/test_ref_pin.py) run in the adapter's own set order."""
    exe = os.path.join(ADAPTER, "adapter_track_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    g = np.load(os.path.join(GOLD, "ref_track_seed%d.npz" % seed))
    import hashlib
    H, W = synth.K_SHAPE
    seq = synth.Sequence((H, W), seed=seed)
    (L0, R0), (L1, R1) = seq.frame(0), seq.frame(1)
    D0, D1 = synth.dense_disparity((H, W), 2 * seed), synth.dense_disparity((H, W), 2 * seed + 1)
    for a, want in zip((L0, R0, L1, R1, D0, D1), g["input_sha"]):
        assert hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest() == str(want), "synthetic inputs changed: regenerate the golden"
    paths = []
    for name, im in (("L0", L0), ("R0", R0), ("L1", L1), ("R1", R1), ("D0", D0), ("D1", D1)):
        p = str(tmp_path / (name + ".raw")); im.tofile(p); paths.append(p)
    bx = str(tmp_path / "boxes.txt"); np.savetxt(bx, g["boxes"].reshape(-1, 4), fmt="%d")
    fp = str(tmp_path / "F.txt"); open(fp, "w").write(" ".join(repr(float(v)) for v in g["F"].reshape(9)))
    out = str(tmp_path / "out.txt")
    subprocess.check_call([exe, str(W), str(H), "500"] + paths + [bx, fp, out])
    f0, mp, kp, row, fpt, pose = [], [], [], {}, [], None
    for line in open(out):
        t = line.split()
        if t[0] == "f0":
            f0.append([np.float32(v) for v in t[2:6]])
        elif t[0] == "map":
            mp.append((int(t[2]), [np.float32(v) for v in t[3:6]]))
        elif t[0] == "kp":
            kp.append(t[1:])
        elif t[0] == "row":
            row[int(t[1])] = (np.float32(t[2]), int(t[3]))
        elif t[0] == "fpt":
            fpt.append([np.float32(v) for v in t[2:6]])
        elif t[0] == "counts":
            counts = list(map(int, t[1:]))
    f0 = np.array(f0, np.float32)

    def same(a, b):
        a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
        return a.shape == b.shape and (a.view(np.uint32) == b.view(np.uint32)).all()

    # first frame: keypoints, keypoints_r (sticky rx included), depthimg at the keypoints
    assert same(f0[:, 0], g["f0_kps"][:, 0]) and same(f0[:, 1], g["f0_kps"][:, 1])
    assert same(f0[:, 2], g["f0_keypoints_r"][:, 0])
    assert same(f0[:, 3], g["f0_depth_at_kp"])
    # createmappoint: same keypoints get a map point, same world positions (set order differs: compare by keypoint index)
    assert counts[2] == int(g["created"]) == len(mp)
    got_pos = {i: np.array(p, np.float32) for i, p in mp}
    assert sorted(got_pos) == sorted(int(i) for i in g["map_idx"])
    for i, pos in zip(g["map_idx"], g["map_worldpos"]):
        assert same(got_pos[int(i)], pos), i
    # find_feature_matches + the box filter of poseEstimation2D_2D: the point lists given to findFundamentalMat
    fpt = np.array(fpt, np.float32).reshape(-1, 4)
    assert same(fpt[:, :2], g["F_p1"]) and same(fpt[:, 2:], g["F_p2"])
    # second frame after the re-extraction: keypoints and descriptors
    assert same([np.float32(r[1]) for r in kp], g["cur_kps"][:, 0]) and same([np.float32(r[2]) for r in kp], g["cur_kps"][:, 1])
    desc = np.array([[int(v) for v in r[4:36]] for r in kp], np.uint8)
    assert (desc == g["cur_desc"]).all()
    # pass 1: match_score for every row with a map point, and the rows the veto turned bad
    has = g["last_has_mp"].astype(bool)
    assert sorted(row) == list(np.nonzero(has)[0])
    for i, (score, bad) in row.items():
        assert np.float32(score).view(np.uint32) == g["cur_match_score"][i].view(np.uint32), i
        assert bad == int(g["last_mp_bad"][i]), i
    assert int(g["last_mp_bad"].sum()) == sum(b for _, b in row.values())
    # final MapPoints: oracle replay with the adapter's own pass-2 order
    live = has.astype(np.uint8)
    rows1 = np.zeros((len(has), 32), np.uint8); rows1[:len(g["f0_desc"])] = g["f0_desc"][:len(has)]
    veto = dict(boxes=g["boxes"], F=g["F"], row_xy=g["f0_kps"][:len(has), :2], cur_xy=g["cur_kps"][:, :2]) if len(g["boxes"]) else None
    p1 = O.match_greedy(rows1, g["cur_desc"], 0, row_live=live, veto=veto)
    order = np.array([i for i, _ in mp], np.int64)
    live2 = np.ones(len(order), np.uint8)
    live2[(p1["row_bad"][order] == 1) | (p1["row_claimed"][order] == 1)] = 0
    p2 = O.match_greedy(g["f0_desc"][order], g["cur_desc"], 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=live2,
                        row_base=10 ** 6)
    expect = np.full(len(kp), -1, np.int64)
    for j in np.nonzero(p2["claim_row"] >= 0)[0]:
        r = int(p2["claim_row"][j])
        expect[j] = r if r < 10 ** 6 else order[r - 10 ** 6]
    assert ([int(r[3]) for r in kp] == expect).all()
    assert (expect >= 0).sum() > 40
    # and the golden's own final assignment is the oracle's in the reference's set order (runs without /root/reference)
    ordg = g["map_idx"].astype(np.int64)
    l2 = np.ones(len(ordg), np.uint8); l2[(p1["row_bad"][ordg] == 1) | (p1["row_claimed"][ordg] == 1)] = 0
    q2 = O.match_greedy(g["map_desc"], g["cur_desc"], 1, claimed=p1["claimed"], claim_row=p1["claim_row"], row_live=l2, row_base=10 ** 6)
    eg = np.full(len(kp), -1, np.int64)
    for j in np.nonzero(q2["claim_row"] >= 0)[0]:
        r = int(q2["claim_row"][j])
        eg[j] = r if r < 10 ** 6 else ordg[r - 10 ** 6]
    assert (eg == g["cur_mp_idx"][:len(kp)]).all()


def test_adapter_refuses_boxes_without_F(tmp_path):
    """offline_box given but no fundamental matrix (ADVICE): the reference would dereference an empty Mat; the
    adapter raises instead of silently skipping the veto."""
    exe = os.path.join(ADAPTER, "adapter_track_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    H, W = synth.K_SHAPE
    seq = synth.Sequence((H, W), seed=5)
    (L0, R0), (L1, R1) = seq.frame(0), seq.frame(1)
    D = synth.dense_disparity((H, W), 10)
    paths = []
    for name, im in (("L0", L0), ("R0", R0), ("L1", L1), ("R1", R1), ("D0", D), ("D1", D)):
        p = str(tmp_path / (name + ".raw")); im.tofile(p); paths.append(p)
    bx = str(tmp_path / "boxes.txt"); open(bx, "w").write("300 700 100 300\n")
    fp = str(tmp_path / "F.txt"); open(fp, "w").write("")
    r = subprocess.run([exe, str(W), str(H), "500"] + paths + [bx, fp, str(tmp_path / "out.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "fundamental" in r.stderr


def test_track_loop_in_cpp_over_the_device_resident_state(tmp_path):
    """adapter/track_loop_test.cc — the host loop of INTEGRATION.md section 3a in plain C++ over the C ABI: six frames of
    one sequence through svo_frame_in.track_seq with boxes and F on some frames.  Its dump (claims, the point every
    keypoint owns, bad rows) must equal oracle/track.py stepped with the keypoints and depths the run itself reports
    (extraction and stereo have their own parity tests; tests/test_oracle_track.py pins the oracle to the reference)."""
    from oracle import track as T
    import svo
    exe = os.path.join(ADAPTER, "track_loop_test")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", ADAPTER, "-s"])
    H, W, NF, NT = 240, 400, 500, 6
    seq = synth.Sequence((H, W), seed=61)
    frames = [seq.frame(t) for t in range(NT)]
    np.concatenate([np.stack(f) for f in frames]).tofile(str(tmp_path / "frames.raw"))
    boxes = {t: [[90, 220, 30, 100], [250, 350, 60, 200]] for t in (1, 3, 4)}
    with open(tmp_path / "boxes.txt", "w") as f:
        for t, bl in boxes.items():
            for b in bl:
                f.write("%d %d %d %d %d\n" % (t, *b))
    out = str(tmp_path / "out.txt")
    subprocess.check_call([exe, str(W), str(H), str(NF), str(NT), str(tmp_path / "frames.raw"), str(tmp_path / "boxes.txt"), out])
    kps = {t: [] for t in range(NT)}; bad = {t: [] for t in range(NT)}; head = {}
    for line in open(out):
        p = line.split()
        if p[0] == "frame":
            head[int(p[1])] = [int(v) for v in p[2:]]
        elif p[0] == "kp":
            kps[int(p[1])].append([np.float32(v) for v in p[3:6]] + [int(p[6]), int(p[7])] + [np.float32(v) for v in p[8:11]])
        elif p[0] == "bad":
            bad[int(p[1])].append(int(p[2]))
    # the descriptors are not in the dump: take them from the same extractor through the Python binding
    ctx = svo.Context(W, H, nfeatures=NF, max_batch=1, lanes=1, max_rows=1000)
    K4 = (707.0912, 707.0912, 601.8873, 183.1104)
    F = np.array([[1.1e-9, 2.3e-7, -3.1e-4], [-2.2e-7, 0.9e-9, 0.8312], [2.9e-4, -0.8297, 1.0]], np.float64)
    trk = T.Tracker(window=4, map_cap=3000)
    total_claims = total_bad = 0
    for t in range(NT):
        kp, desc = ctx.extract(frames[t][0])
        rows = kps[t]
        n_left, n_prev, n_map, _ = head[t]
        assert n_left == len(kp) == len(rows)
        xy = np.array([[r[0], r[1]] for r in rows], np.float32); depth = np.array([r[2] for r in rows], np.float32)
        assert (xy[:, 0] == kp["x"]).all() and (xy[:, 1] == kp["y"]).all()
        o = trk.step(xy, desc, depth, t, boxes=boxes.get(t), F=F if t in boxes else None, K4=K4)
        assert (n_prev, n_map) == (o["n_prev"], o["n_map"]), t
        assert [r[3] for r in rows] == o["claim_row"].tolist(), t
        assert [r[4] for r in rows] == o["mp_create"].tolist(), t
        got_xyz = np.array([r[5:8] for r in rows], np.float32)
        assert (got_xyz.view(np.uint32) == o["mp_xyz"].view(np.uint32)).all(), t
        if o["n_prev"]:
            assert bad[t] == np.nonzero(o["p1_row_bad"])[0].tolist(), t
            total_bad += len(bad[t])
        total_claims += int((o["claim_row"] >= 0).sum())
    ctx.close()
    assert total_claims > 300
