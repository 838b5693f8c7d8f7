"""ctypes view of oracle/libsvo_oracle.so — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module (see oracle/svo_oracle.c).
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
MAX_LEVELS = 16

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


class Pyramid(C.Structure):
    _fields_ = [("nlevels", C.c_int), ("w", C.c_int * MAX_LEVELS), ("h", C.c_int * MAX_LEVELS),
                ("scale", C.c_float * MAX_LEVELS),
                ("img", C.POINTER(C.c_uint8) * MAX_LEVELS), ("blur", C.POINTER(C.c_uint8) * MAX_LEVELS)]

    def level(self, l, blurred=False):
        p = (self.blur if blurred else self.img)[l]
        return np.ctypeslib.as_array(p, shape=(self.h[l], self.w[l])).copy()


def build(force=False):
    so = os.path.join(_HERE, "libsvo_oracle.so")
    if force or not os.path.exists(so):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.svo_o_fast_atan2.restype = C.c_float
        _LIB.svo_o_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _LIB


_NATIVE = None


def native_matchers():
    """bench.py's CPU arm only: the same svo_matchers.c built ON THE RUNNING HOST with -O3 -march=native (what the
    reference's CMakeLists.txt:10-11 passes), used for the integer matchers (BF, pass 1, pass 2).  The float stages keep
    the portable -O2 -ffp-contract=off build, whose bits are what the parity tests pin.  Built into oracle/_native/
    (git-ignored) and keyed by the CPU model so a snapshot built elsewhere is never run on a CPU it was not built for."""
    global _NATIVE
    if _NATIVE is None:
        import hashlib
        try:
            model = [ln for ln in open("/proc/cpuinfo") if ln.startswith(("model name", "flags"))][:2]
        except OSError:
            model = []
        tag = hashlib.sha1("".join(model).encode()).hexdigest()[:10]
        d = os.path.join(_HERE, "_native")
        so = os.path.join(d, "libsvo_matchers_%s.so" % tag)
        if not os.path.exists(so):
            os.makedirs(d, exist_ok=True)
            subprocess.check_call(["gcc", "-O3", "-march=native", "-fPIC", "-shared", "-std=c11", "-D_GNU_SOURCE",
                                   os.path.join(_HERE, "svo_matchers.c"),
                                   "-o", so + ".tmp", "-lm"])
            os.replace(so + ".tmp", so)
        _NATIVE = C.CDLL(so)
    return _NATIVE


def pyramid_from_levels(levels, scales):
    """A Pyramid over caller-built un-blurred levels (contiguous u8 arrays), e.g. a chained
    cv2.resize(INTER_LINEAR_EXACT) — identical to the oracle's own levels (tests/test_oracle_vs_cv2.py).
    The arrays must outlive the struct (they are kept on it)."""
    pyr = Pyramid()
    pyr.nlevels = len(levels)
    keep = []
    for l, (a, s) in enumerate(zip(levels, scales)):
        a = np.ascontiguousarray(a, np.uint8)
        keep.append(a)
        pyr.w[l], pyr.h[l], pyr.scale[l] = a.shape[1], a.shape[0], float(s)
        pyr.img[l] = a.ctypes.data_as(C.POINTER(C.c_uint8))
    pyr._keep = keep
    return pyr


def _p(a, t=C.c_void_p):
    return a.ctypes.data_as(t) if a is not None else None


def geometry(W, H, nlevels=8, scale=1.2, nfeatures=500):
    lw = np.zeros(nlevels, np.int32); lh = np.zeros(nlevels, np.int32)
    ls = np.zeros(nlevels, np.float32); q = np.zeros(nlevels, np.int32)
    lib().svo_o_geometry(W, H, nlevels, C.c_float(scale), nfeatures, _p(lw), _p(lh), _p(ls), _p(q))
    return lw, lh, ls, q


def resize(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.empty((dh, dw), np.uint8)
    lib().svo_o_resize(_p(src), src.shape[1], src.shape[0], src.shape[1], _p(dst), dw, dh, dw)
    return dst


def fast_nms(img, threshold=20, border=31):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = ((w + 1) // 2) * ((h + 1) // 2)
    xs = np.empty(cap, np.int32); ys = np.empty(cap, np.int32); sc = np.empty(cap, np.int32)
    n = lib().svo_o_fast_nms(_p(img), w, h, w, threshold, border, _p(xs), _p(ys), _p(sc), cap)
    return xs[:n].copy(), ys[:n].copy(), sc[:n].copy()


def retain_best(resp, n_points):
    resp = np.array(resp, np.float32)
    idx = np.arange(resp.size, dtype=np.int32)
    k = lib().svo_o_retain_best(_p(resp), _p(idx), resp.size, n_points)
    return idx[:k].copy(), resp[:k].copy()


def harris(img, xs, ys):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    r = np.empty(xs.size, np.float32)
    lib().svo_o_harris(_p(img), img.shape[1], _p(xs), _p(ys), xs.size, _p(r))
    return r


def ic_angle(img, xs, ys):
    img = np.ascontiguousarray(img, np.uint8)
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    r = np.empty(xs.size, np.float32)
    lib().svo_o_ic_angle(_p(img), img.shape[1], _p(xs), _p(ys), xs.size, _p(r))
    return r


def blur7(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().svo_o_blur7(_p(img), img.shape[1], img.shape[0], img.shape[1], _p(out), img.shape[1])
    return out


def orb(gray, nfeatures=500, scale=1.2, nlevels=8, fast_threshold=20, with_pyramid=False, distribution=0):
    """cv::ORB::detectAndCompute restatement -> (keypoints[KP_DTYPE], descriptors[n,32], Pyramid|None).
    distribution=1: the opt-in quadtree distribution (svo_octree_oracle.c) instead of the two retainBest culls."""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    cap = 4 * nfeatures + 4096
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    pyr = Pyramid() if with_pyramid else None
    n = lib().svo_o_orb_ex(_p(gray), W, H, W, nfeatures, C.c_float(scale), nlevels, fast_threshold, distribution,
                           _p(kps), _p(desc), cap, C.byref(pyr) if pyr is not None else None)
    assert 0 <= n <= cap, n
    return kps[:n].copy(), desc[:n].copy(), pyr


def distribute_octree(xs, ys, score, rect, N):
    """Quadtree distribution of raster-ordered points inside rect = (x0, y0, x1, y1) -> kept indices (node order)."""
    xs = np.ascontiguousarray(xs, np.int32); ys = np.ascontiguousarray(ys, np.int32)
    score = np.ascontiguousarray(score, np.int32)
    out = np.zeros(max(len(xs), 1), np.int32)
    k = lib().svo_o_distribute_octree(_p(xs), _p(ys), _p(score), len(xs), rect[0], rect[1], rect[2], rect[3], N, _p(out))
    return out[:k].copy()


def pyramid_free(pyr):
    lib().svo_o_pyramid_free(C.byref(pyr))


def hamming(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().svo_o_hamming(_p(a), _p(b))


def match_bf(q, t, L=None):
    q = np.ascontiguousarray(q, np.uint8).reshape(-1, 32); t = np.ascontiguousarray(t, np.uint8).reshape(-1, 32)
    idx = np.empty(len(q), np.int32); dist = np.empty(len(q), np.int32); keep = np.empty(len(q), np.uint8)
    (L or lib()).svo_o_match_bf(_p(q), len(q), _p(t), len(t), _p(idx), _p(dist), _p(keep))
    return idx, dist, keep


def match_greedy(rows, cur, mode, claimed=None, row_live=None, row_base=0, claim_row=None,
                 win_uvr=None, cur_xy=None, veto=None, L=None):
    """veto (pass 1): dict(boxes (n,4) int32 left/right/top/bottom, F (3,3) f64, row_xy (M,2), cur_xy (N,2)) —
    the YOLO-box + epipolar "dynamic" test of src/pnpmatch.cc:103-144; row_bad marks map points turned bad."""
    rows = np.ascontiguousarray(rows, np.uint8).reshape(-1, 32); cur = np.ascontiguousarray(cur, np.uint8).reshape(-1, 32)
    M, N = len(rows), len(cur)
    claimed = np.zeros(N, np.uint8) if claimed is None else np.ascontiguousarray(claimed, np.uint8).copy()
    claim_row = np.full(N, -1, np.int32) if claim_row is None else np.ascontiguousarray(claim_row, np.int32).copy()
    if row_live is not None:
        row_live = np.ascontiguousarray(row_live, np.uint8)
    if win_uvr is not None:
        win_uvr = np.ascontiguousarray(win_uvr, np.float32); cur_xy = np.ascontiguousarray(cur_xy, np.float32)
    bi = np.empty(M, np.int32); b = np.empty(M, np.int32); s = np.empty(M, np.int32); rc = np.empty(M, np.uint8)
    bad = np.zeros(M, np.uint8)
    boxes = F = rxy = vxy = None
    if veto is not None:
        boxes = np.ascontiguousarray(veto["boxes"], np.int32).reshape(-1, 4)
        F = np.ascontiguousarray(veto["F"], np.float64).reshape(9)
        rxy = np.ascontiguousarray(veto["row_xy"], np.float32).reshape(-1, 2)
        vxy = np.ascontiguousarray(veto["cur_xy"], np.float32).reshape(-1, 2)
        assert len(rxy) == M and len(vxy) == N
    (L or lib()).svo_o_match_greedy_veto(_p(rows), M, _p(cur), N, mode, _p(row_live), _p(claimed), _p(claim_row), row_base,
                                  _p(bi), _p(b), _p(s), _p(rc), _p(win_uvr), _p(cur_xy),
                                  _p(boxes), 0 if boxes is None else len(boxes), _p(F), _p(rxy), _p(vxy), _p(bad))
    return dict(best_idx=bi, best=b, second=s, row_claimed=rc, claimed=claimed, claim_row=claim_row, row_bad=bad)


def project_map(xyz, octave, Tcw, K4, W, H, th=7.0, nlevels=8, scale=1.2):
    """Projection windows (u, v, r) of map points under the predicted pose (opt-in pass-2 mode; ORB-SLAM2 SearchByProjection)."""
    xyz = np.ascontiguousarray(xyz, np.float32).reshape(-1, 3)
    octave = np.ascontiguousarray(octave, np.int32)
    T = np.ascontiguousarray(Tcw, np.float32).reshape(16)
    _, _, ls, _ = geometry(W, H, nlevels, scale, 500)
    out = np.zeros((len(xyz), 3), np.float32)
    fx, fy, cx, cy = [C.c_float(float(v)) for v in K4]
    lib().svo_o_project_map(_p(xyz), _p(octave), len(xyz), _p(T), fx, fy, cx, cy, W, H, C.c_float(th), _p(ls), nlevels, _p(out))
    return out


def disp2depth(disp, bf):
    disp = np.ascontiguousarray(disp, np.float32)
    out = np.empty_like(disp)
    lib().svo_o_disp2depth(_p(disp), _p(out), C.c_size_t(disp.size), C.c_float(bf))
    return out


def stereo_sparse(kl, dl, pl, kr, dr, pr, bf, b):
    kl = np.ascontiguousarray(kl); kr = np.ascontiguousarray(kr)
    dl = np.ascontiguousarray(dl, np.uint8); dr = np.ascontiguousarray(dr, np.uint8)
    n = len(kl)
    ur = np.empty(n, np.float32); dep = np.empty(n, np.float32)
    mr = np.empty(n, np.int32); sad = np.empty(n, np.int32)
    lib().svo_o_stereo_sparse(_p(kl), _p(dl), n, _p(kr), _p(dr), len(kr), C.byref(pl), C.byref(pr),
                              C.c_float(bf), C.c_float(b), _p(ur), _p(dep), _p(mr), _p(sad))
    return ur, dep, mr, sad


def pose_optimize(Xw, obs, K4, Tcw, iterations=10):
    """Optimizer::PoseOptimization restatement -> (Tcw_out[4,4] f32, outer iterations, robust chi2)."""
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
    Tin = np.ascontiguousarray(Tcw, np.float32).reshape(4, 4); Tout = np.empty((4, 4), np.float32)
    stats = np.zeros(2, np.float64)
    fx, fy, cx, cy = [C.c_float(float(v)) for v in K4]
    L = lib(); L.svo_o_pose_optimize.restype = C.c_int
    L.svo_o_pose_optimize(_p(Xw), _p(obs), len(Xw), fx, fy, cx, cy, _p(Tin), _p(Tout), iterations, _p(stats))
    return Tout, int(stats[0]), float(stats[1])


def pnp_ransac(pts3d, pts2d, K4, iterations=100, reproj_err=8.0, seed=1, refine_iters=10):
    """P3P RANSAC stand-in for cv::solvePnPRansac -> (n_inliers, R[3,3] f64, t[3] f64, mask[n] u8, info[3])."""
    p3 = np.ascontiguousarray(pts3d, np.float32).reshape(-1, 3); p2 = np.ascontiguousarray(pts2d, np.float32).reshape(-1, 2)
    R = np.zeros((3, 3), np.float64); t = np.zeros(3, np.float64)
    mask = np.zeros(len(p3), np.uint8); info = np.zeros(3, np.int32)
    fx, fy, cx, cy = [C.c_float(float(v)) for v in K4]
    L = lib(); L.svo_o_pnp_ransac.restype = C.c_int
    n = L.svo_o_pnp_ransac(_p(p3), _p(p2), len(p3), fx, fy, cx, cy, iterations, C.c_float(reproj_err),
                           C.c_uint32(seed), refine_iters, _p(R), _p(t), _p(mask), _p(info))
    return n, R, t, mask, info


def bgr2gray(bgr):
    """cv::cvtColor(BGR2GRAY) restatement (what cv::ORB applies to colour input)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w), np.uint8)
    lib().svo_o_bgr2gray(_p(bgr), w, h, bgr.strides[0], _p(out), w)
    return out
