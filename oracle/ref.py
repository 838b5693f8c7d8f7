"""ctypes driver of oracle/_ref/libsvo_ref.so — TEST INFRASTRUCTURE ONLY.

libsvo_ref.so is the reference's own src/pnpmatch.cc, src/frame.cc and src/mappoint.cc compiled UNMODIFIED from
/root/reference (recipe: `make -C oracle ref`) against oracle/ref_stubs/minicv.hpp.  The OpenCV calls those files make
(cv::ORB, findFundamentalMat, solvePnPRansac, Rodrigues) are forwarded to the real cv2 through the callbacks below,
with the reference's literal arguments.  Used by tests/test_ref_pin.py to pin oracle/svo_oracle.c's matchers, and by
tests/golden/make_golden_ref.py to record what the reference computes for the drop-in adapter's GPU tests.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libsvo_ref.so")
REFERENCE = "/root/reference"
_LIB = None
_KEEP = []

ORB_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                     C.POINTER(C.c_uint8), C.c_int, C.c_int)
FUND_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_double))
PNP_FN = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                     C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_int)
ROD_FN = C.CFUNCTYPE(None, C.POINTER(C.c_double), C.POINTER(C.c_double))

LOG = {"fundamental": [], "pnp": []}     # what the hooks saw / returned, for the tests
# Tests that need hand-made feature sets (adversarial descriptor chains, ties, thresholds) set this to a callable
# (img, what) -> (n x 7 float32 keypoints, n x 32 uint8 descriptors) that answers the reference's cv::ORB calls instead
# of cv2; what = 0 detectAndCompute, 1 detect, 2 compute (the keypoints handed in are the ones it returned for 1).
ORB_OVERRIDE = None


def available():
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "src"))


def build(force=False):
    if force or not os.path.exists(SO):
        subprocess.check_call(["make", "-C", _HERE, "-s", "ref"])
    return SO


def _image(ptr, rows, cols, step, ch):
    buf = (C.c_uint8 * (rows * step)).from_address(ptr)
    a = np.frombuffer(buf, np.uint8).reshape(rows, step)[:, :cols * ch]
    return a.reshape(rows, cols, ch).copy() if ch > 1 else a.copy()


def _orb_hook(ptr, rows, cols, step, ch, what, kps, desc, cap, n_in):
    import cv2
    img = _image(ptr, rows, cols, step, ch)
    if ORB_OVERRIDE is not None:
        k7, d = ORB_OVERRIDE(img, what)
        n = len(k7)
        assert n <= cap and (what != 2 or n == n_in)
        flat = np.ascontiguousarray(k7, np.float32).reshape(-1)
        for i in range(7 * n):
            kps[i] = flat[i]
        if what != 1 and n:
            C.memmove(desc, np.ascontiguousarray(d, np.uint8).ctypes.data, n * 32)
        return n
    orb = cv2.ORB_create()                       # cv::ORB::create(): src/frame.cc:77, src/pnpmatch.cc:261-262
    if what == 0:
        k, d = orb.detectAndCompute(img, None)
    elif what == 1:
        k, d = orb.detect(img, None), None
    else:
        kin = [cv2.KeyPoint(kps[7 * i], kps[7 * i + 1], kps[7 * i + 2], kps[7 * i + 3], kps[7 * i + 4],
                            int(kps[7 * i + 5]), int(kps[7 * i + 6])) for i in range(n_in)]
        k, d = orb.compute(img, kin)
    n = len(k)
    assert n <= cap
    for i, p in enumerate(k):
        kps[7 * i], kps[7 * i + 1], kps[7 * i + 2], kps[7 * i + 3] = p.pt[0], p.pt[1], p.size, p.angle
        kps[7 * i + 4], kps[7 * i + 5], kps[7 * i + 6] = p.response, float(p.octave), float(p.class_id)
    if d is not None and n:
        C.memmove(desc, np.ascontiguousarray(d).ctypes.data, n * 32)
    return n


def _fund_hook(p1, p2, n, F9):
    import cv2
    a = np.ctypeslib.as_array(p1, (n, 2)).copy() if n else np.zeros((0, 2), np.float32)
    b = np.ctypeslib.as_array(p2, (n, 2)).copy() if n else np.zeros((0, 2), np.float32)
    F = None
    if n >= 8:
        F, _ = cv2.findFundamentalMat(a, b, cv2.FM_8POINT)     # src/pnpmatch.cc:336
    LOG["fundamental"].append(dict(p1=a, p2=b, F=None if F is None else F.copy()))
    if F is None or F.shape != (3, 3):
        return 0
    for i in range(9):
        F9[i] = float(F.reshape(9)[i])
    return 1


def _pnp_hook(p3, p2, n, K9, rvec, tvec, inl, cap):
    import cv2
    a = np.ctypeslib.as_array(p3, (n, 3)).copy() if n else np.zeros((0, 3), np.float32)
    b = np.ctypeslib.as_array(p2, (n, 2)).copy() if n else np.zeros((0, 2), np.float32)
    K = np.ctypeslib.as_array(K9, (3, 3)).copy()
    rec = dict(p3=a, p2=b, ok=False)
    LOG["pnp"].append(rec)
    if n < 4:
        return 0
    cv2.setRNGSeed(0)
    try:   # cv::solvePnPRansac(pts3d, pts2d, K, Mat(), rvec, tvec, false, 100, 8.0, 0.99, inliers)  src/pnpmatch.cc:227
        ok, r, t, inliers = cv2.solvePnPRansac(a, b, K, None, None, None, False, 100, 8.0, 0.99)
    except cv2.error:
        return 0
    if not ok or inliers is None:
        return 0
    for i in range(3):
        rvec[i] = float(r[i, 0]); tvec[i] = float(t[i, 0])
    m = min(len(inliers), cap)
    for i in range(m):
        inl[i] = int(inliers[i, 0])
    rec.update(ok=True, rvec=r.copy(), tvec=t.copy(), inliers=inliers[:, 0].copy())
    return m


def _rod_hook(rvec, R9):
    import cv2
    R, _ = cv2.Rodrigues(np.array([rvec[0], rvec[1], rvec[2]], np.float64))   # src/pnpmatch.cc:238
    for i in range(9):
        R9[i] = float(R.reshape(9)[i])


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        hooks = (ORB_FN(_orb_hook), FUND_FN(_fund_hook), PNP_FN(_pnp_hook), ROD_FN(_rod_hook))
        _KEEP.extend(hooks)
        L.ref_set_hooks(*hooks)
        L._svo_hooks = hooks
        L.ref_frame_new.restype = C.c_void_p
        L.ref_frame_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float, C.c_void_p,
                                    C.c_int, C.c_double, C.c_long]
        L.ref_frame_copy.restype = C.c_void_p
        L.ref_frame_copy.argtypes = [C.c_void_p]
        for name in ("ref_frame_free", "ref_frame_featuredetect", "ref_frame_stereo"):
            getattr(L, name).argtypes = [C.c_void_p]
            getattr(L, name).restype = None
        L.ref_frame_set_disp.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_frame_set_pose.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_frame_unproject.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_void_p, C.c_void_p]
        L.ref_localmap_new.restype = C.c_void_p
        L.ref_localmap_size.argtypes = [C.c_void_p]
        L.ref_localmap_list.argtypes = [C.c_void_p] * 6 + [C.c_int]
        L.ref_localmap_age.argtypes = [C.c_void_p, C.c_int]
        L.ref_frame_createmappoint.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_pose_estimation_pnp.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_frame_counts.argtypes = [C.c_void_p] * 4
        L.ref_frame_get.argtypes = [C.c_void_p] * 10
        L.ref_frame_get_images.argtypes = [C.c_void_p] * 3
        L.ref_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def activate():
    """(Re-)install THIS module instance's hooks.  oracle/ref_g2o.py binds a second instance of this module to another
    library; if the two libraries ever shared their hook table (a GNU_UNIQUE symbol — the recipe builds with
    -fno-gnu-unique so that they do not) the instance loaded last would answer both, and its LOG would get the records."""
    L = lib()
    L.ref_set_hooks(*L._svo_hooks)
    return L


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().ref_descriptor_distance(_p(a), _p(b))


class LocalMap:
    """std::set<mappoint*> localmappoints (src/Tracking.cc)."""

    def __init__(self):
        self.h = lib().ref_localmap_new()

    def __len__(self):
        return lib().ref_localmap_size(self.h)

    def age(self, frame_num):
        """The 4-frame window of Tracking::Track (src/Tracking.cc:239-250)."""
        return lib().ref_localmap_age(self.h, int(frame_num))

    def list(self):
        """Map points in the set's own iteration order (what pass 2 walks)."""
        n = len(self)
        cid = np.zeros(n, np.int32); idx = np.zeros(n, np.int32); bad = np.zeros(n, np.uint8)
        pos = np.zeros((n, 3), np.float32); desc = np.zeros((n, 32), np.uint8)
        k = lib().ref_localmap_list(self.h, _p(cid), _p(idx), _p(bad), _p(pos), _p(desc), n)
        assert k == n
        return dict(create_id=cid, idx=idx, bad=bad, worldpos=pos, desc=desc)


class Frame:
    """The reference's `frame` (include/frame.h), driven like Tracking::Track drives it."""

    def __init__(self, left=None, right=None, K=None, bf=None, boxes=(), ts=0.0, fid=0, handle=None):
        L = lib()
        if handle is not None:
            self.h = handle
            return
        left = np.ascontiguousarray(left, np.uint8); right = np.ascontiguousarray(right, np.uint8)
        h, w = left.shape[:2]
        ch = 1 if left.ndim == 2 else left.shape[2]
        K9 = np.ascontiguousarray(K, np.float32).reshape(9)
        bx = np.ascontiguousarray(np.asarray(boxes, np.int32).reshape(-1, 4))
        self.shape = (h, w)
        self.h = L.ref_frame_new(_p(left), _p(right), w, h, ch, _p(K9), float(bf), _p(bx), len(bx), float(ts), int(fid))

    def copy(self):
        f = Frame(handle=lib().ref_frame_copy(self.h))
        f.shape = self.shape
        return f

    def featuredetect(self):
        lib().ref_frame_featuredetect(self.h)

    def set_disp(self, disp):
        disp = np.ascontiguousarray(disp, np.float32)
        assert disp.shape == self.shape
        lib().ref_frame_set_disp(self.h, _p(disp))

    def stereo(self):
        """computekeypoint_r(); disp2Depth(bf)  (src/Tracking.cc:227-228)."""
        lib().ref_frame_stereo(self.h)

    def set_pose(self, T):
        T = np.ascontiguousarray(T, np.float32).reshape(16)
        lib().ref_frame_set_pose(self.h, _p(T))

    def unproject(self, u, v, z):
        out = np.zeros(3, np.float32); ok = C.c_int(0)
        lib().ref_frame_unproject(self.h, float(u), float(v), float(z), _p(out), C.byref(ok))
        return out if ok.value else None

    def createmappoint(self, localmap):
        return lib().ref_frame_createmappoint(self.h, localmap.h)

    def pose_estimation_pnp(self, last, localmap, K):
        K9 = np.ascontiguousarray(K, np.float32).reshape(9)
        return lib().ref_pose_estimation_pnp(self.h, last.h, localmap.h, _p(K9))

    def state(self):
        N = C.c_int(); nk = C.c_int(); nr = C.c_int()
        nd = lib().ref_frame_counts(self.h, C.byref(N), C.byref(nk), C.byref(nr))
        N, nk, nr = N.value, nk.value, nr.value
        kps = np.zeros((nk, 6), np.float32); desc = np.zeros((nd, 32), np.uint8); kpr = np.zeros((nr, 2), np.float32)
        dep = np.full(nk, np.nan, np.float32); ms = np.zeros(N, np.float32)
        cid = np.zeros(N, np.int32); idx = np.zeros(N, np.int32); bad = np.zeros(N, np.uint8); T = np.zeros(16, np.float32)
        lib().ref_frame_get(self.h, _p(kps), _p(desc), _p(kpr), _p(dep), _p(ms), _p(cid), _p(idx), _p(bad), _p(T))
        return dict(N=N, kps=kps, desc=desc, keypoints_r=kpr, depth_at_kp=dep, match_score=ms, mp_create_id=cid, mp_idx=idx,
                    mp_bad=bad, Tcw=T.reshape(4, 4))

    def images(self):
        h, w = self.shape
        disp = np.zeros((h, w), np.float32); depth = np.zeros((h, w), np.float32)
        lib().ref_frame_get_images(self.h, _p(disp), _p(depth))
        return disp, depth


def run_two_frames(frames, disps, K, bf, boxes, pose0=None):
    """Two stereo frames through the reference exactly as Tracking::Track / Tracklastframe drive it
    (src/Tracking.cc:184, :225-238, :114).  frames = ((L0, R0), (L1, R1)); disps = the dense disparity images standing
    in for frame::MB's output.  Returns everything the pin tests and the golden fixtures compare."""
    (L0, R0), (L1, R1) = frames
    activate()
    LOG["fundamental"].clear(); LOG["pnp"].clear()
    f0 = Frame(L0, R0, K, bf, boxes, 0.0, 0)
    if pose0 is not None:
        f0.set_pose(pose0)
    f0.featuredetect(); f0.set_disp(disps[0]); f0.stereo()
    s0 = f0.state()
    last = f0.copy()                                  # Tracking.cc:237
    lm = LocalMap()
    created = last.createmappoint(lm)                 # :238
    before = dict(last=last.state(), map=lm.list())
    f1 = Frame(L1, R1, K, bf, boxes, 0.1, 1)
    f1.featuredetect(); f1.set_disp(disps[1]); f1.stereo()
    f1.pose_estimation_pnp(last, lm, K)               # Tracking.cc:114
    d0, z0 = f0.images()
    return dict(f0=s0, created=created, before=before, cur=f1.state(), last=last.state(), map=lm.list(),
                F=dict(LOG["fundamental"][-1]), pnp=dict(LOG["pnp"][-1]), disp0=d0, depth0=z0)


def run_sequence(frames, disps, K, bf, boxes_per_frame):
    """n stereo frames through the reference as Tracking::Track drives it (src/Tracking.cc:184, :225-250): per frame
    featuredetect, the dense disparity standing in for frame::MB, computekeypoint_r, disp2Depth, poseEstimationPnP against
    the last frame and the local map (frame 0: the map points Tracking::init / createmappoint make), then
    lastframe = frame(currentframe), createmappoint and the 4-frame window.  Optimizer::PoseOptimization (g2o, not built
    here) only refines the pose and is left out: nothing on the matching path reads it.
    Returns one record per frame: the current frame's state after the matching, the fundamental matrix it used, the last
    frame's state and the local map (in the set's own order) after createmappoint + ageing."""
    activate()
    lm = LocalMap()
    last = None
    out = []
    for t, ((L, R), disp) in enumerate(zip(frames, disps)):
        LOG["fundamental"].clear()
        cur = Frame(L, R, K, bf, boxes_per_frame[t], 0.1 * t, t)
        cur.featuredetect(); cur.set_disp(disp); cur.stereo()
        map_before = lm.list()
        if last is not None:
            cur.pose_estimation_pnp(last, lm, K)                  # Tracking.cc:114
        rec = dict(cur=cur.state(), map_before=map_before, last_after_match=None if last is None else last.state(),
                   F=dict(LOG["fundamental"][-1])["F"] if LOG["fundamental"] else None)
        last = cur.copy()                                         # :237
        rec["created"] = last.createmappoint(lm)                  # :238
        rec["erased"] = lm.age(t)                                 # :239-250
        rec["last"] = last.state(); rec["map"] = lm.list()
        out.append(rec)
    return out
