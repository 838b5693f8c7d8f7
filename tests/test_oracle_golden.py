"""Pin the CPU oracle against the committed OpenCV golden vectors (no cv2, no GPU needed)."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
ORB_CASES = sorted(p for p in glob.glob(os.path.join(G, "*seed*.npz")) if not os.path.basename(p).startswith("ref_track"))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("path", ORB_CASES, ids=[os.path.basename(p)[:-4] for p in ORB_CASES])
def test_orb_bit_exact_vs_cv2_golden(path):
    g = np.load(path)
    img = g["image"]
    assert hashlib.sha256(img.tobytes()).hexdigest() == str(g["image_sha256"])
    kp, desc, _ = O.orb(img, int(g["nfeatures"]))
    ref = g["kp"]
    assert len(kp) == len(ref)
    for f in ("x", "y", "size", "angle", "response"):
        assert (bits(kp[f]) == bits(ref[f])).all(), f
    assert (kp["octave"] == ref["octave"]).all()
    assert (desc == g["desc"]).all()


@pytest.mark.parametrize("path", ORB_CASES, ids=[os.path.basename(p)[:-4] for p in ORB_CASES])
def test_stages_vs_cv2_golden(path):
    g = np.load(path)
    img = g["image"]
    l1 = g["level1"]
    assert (O.resize(img, l1.shape[1], l1.shape[0]) == l1).all()
    xs, ys, sc = O.fast_nms(img, 20, border=4)
    f = g["fast"]
    keep = (f[:, 0] >= 4) & (f[:, 0] < img.shape[1] - 4) & (f[:, 1] >= 4) & (f[:, 1] < img.shape[0] - 4)
    f = f[keep]
    assert len(f) == len(xs)
    assert (f[:, 0] == xs).all() and (f[:, 1] == ys).all() and (f[:, 2] == sc).all()
    assert (O.blur7(img) == g["blur"]).all()


def test_geometry_matches_survey_tables():
    lw, lh, ls, q = O.geometry(1241, 376, 8, 1.2, 2000)
    assert list(lw) == [1241, 1034, 862, 718, 598, 499, 416, 346]
    assert list(lh) == [376, 313, 261, 218, 181, 151, 126, 105]
    assert list(q) == [434, 362, 302, 251, 209, 175, 145, 122]
    lw, lh, ls, q = O.geometry(2560, 720, 8, 1.2, 8000)
    assert list(lw) == [2560, 2133, 1778, 1481, 1235, 1029, 857, 714]
    assert list(q) == [1737, 1448, 1207, 1005, 838, 698, 582, 485]
    assert list(O.geometry(1241, 376, 8, 1.2, 4000)[3]) == [869, 724, 603, 503, 419, 349, 291, 242]


def test_bfmatch_vs_cv2_golden():
    g = np.load(os.path.join(G, "bfmatch.npz"))
    idx, dist, keep = O.match_bf(g["q"], g["t"])
    assert (idx == g["train"]).all()
    assert (dist.astype(np.float32) == g["dist"]).all()
    thr = max(2.0 * dist.min(), 30.0)
    assert (keep.astype(bool) == (dist <= thr)).all()
