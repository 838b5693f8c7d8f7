"""ctypes driver of oracle/_ref/libsvo_ref_g2o.so — TEST INFRASTRUCTURE ONLY.

libsvo_ref_g2o.so is the reference's own pose optimisation: src/Optimizer.cc, src/convert.cc and the vendored g2o
(Thirdparty/g2o/g2o/{core,types,stuff}) compiled UNMODIFIED from /root/reference (recipe: `make -C oracle ref_g2o`) against
oracle/ref_stubs_g2o/minieigen.hpp — Eigen is not installed in this image; the stand-in implements the dense matrices, views,
Cholesky and quaternion those sources use — and oracle/ref_stubs/minicv.hpp.  Used by tests/test_ref_pin_pose.py to pin
oracle/svo_pose_oracle.c, and by tests/golden/make_golden_ref_pose.py to record what the reference computes for the GPU box.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libsvo_ref_g2o.so")
REFERENCE = "/root/reference"
MAX_POINTS = 500          # frame::N is fixed (src/frame.cc:54)
_LIB = None


def available():
    return os.path.exists(SO) or os.path.isdir(os.path.join(REFERENCE, "Thirdparty", "g2o"))


def lib():
    global _LIB
    if _LIB is None:
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref_g2o"])
        _LIB = C.CDLL(SO)
        _LIB.ref_pose_optimize.restype = C.c_int
        _LIB.ref_pose_optimize.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 5
    return _LIB


def pose_optimize(Xw, obs, K4, Tcw, has=None):
    """Optimizer::PoseOptimization(frame) of the reference -> (Tcw_out[4,4] f32, number of correspondences).
    Xw: n x 3 map points, obs: n x 2 keypoints, K4 = (fx, fy, cx, cy), has: which keypoints own a map point (default all)."""
    Xw = np.ascontiguousarray(Xw, np.float32).reshape(-1, 3); obs = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
    n = len(Xw)
    assert n <= MAX_POINTS and len(obs) == n
    fx, fy, cx, cy = [float(v) for v in K4]
    K9 = np.array([fx, 0, cx, 0, fy, cy, 0, 0, 1], np.float32)
    hs = np.ones(n, np.uint8) if has is None else np.ascontiguousarray(has, np.uint8)
    Tin = np.ascontiguousarray(Tcw, np.float32).reshape(4, 4); Tout = np.zeros((4, 4), np.float32)
    r = lib().ref_pose_optimize(Tin.ctypes.data, n, obs.ctypes.data, Xw.ctypes.data, hs.ctypes.data, K9.ctypes.data, Tout.ctypes.data)
    assert r >= 0
    return Tout, r


# ---- Tracking::Track itself ------------------------------------------------------------------------------------------
# The same library also holds src/Tracking.cc (compiled unmodified; Pangolin and the viewer are stubbed, frame::MB's
# dense solver hands back the disparity image deposited for the frame) and oracle/ref_harness.cc, so the frame /
# local-map readers of oracle/ref.py work on it: a second instance of that module is bound to this library.
_RR = None


def _ref_module():
    global _RR
    if _RR is None:
        import importlib.util
        spec = importlib.util.spec_from_file_location("oracle._ref_on_g2o", os.path.join(_HERE, "ref.py"))
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        if not os.path.exists(SO):
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref_g2o"])
        m.SO = SO
        m.build = lambda force=False: SO
        L = m.lib()                                   # dlopen + the cv2 hooks (ORB, findFundamentalMat, solvePnPRansac, Rodrigues)
        L.ref_tracking_new.restype = C.c_void_p
        L.ref_tracking_new.argtypes = [C.c_char_p]
        L.ref_tracking_free.argtypes = [C.c_void_p]
        L.ref_tracking_track.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_float, C.c_double, C.c_void_p, C.c_int]
        L.ref_tracking_track.restype = None
        for name in ("ref_tracking_current", "ref_tracking_last"):
            getattr(L, name).restype = C.c_void_p
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_tracking_localmap.restype = C.c_void_p
        L.ref_tracking_frame_num.restype = C.c_int
        _RR = m
    return _RR


def run_tracking(frames, disps, K, bf, boxes_per_frame, tmpdir):
    """n stereo frames through the reference's OWN Tracking::Track (src/Tracking.cc:180-252): new frame, featuredetect,
    MB (the deposited disparity), computekeypoint_r, disp2Depth, Tracklastframe (init / poseEstimationPnP, then
    Optimizer::PoseOptimization), GetVelocity, lastframe = frame(currentframe), createmappoint, the 4-frame window.
    Returns records shaped like oracle/ref.py:run_sequence's (which restates that loop in its harness)."""
    RR = _ref_module()
    L = RR.activate()
    Kf = np.ascontiguousarray(K, np.float32).reshape(3, 3)
    path = os.path.join(str(tmpdir), "settings.yaml")
    with open(path, "w") as f:
        f.write("%%YAML:1.0\nCamera.fx: %.9g\nCamera.fy: %.9g\nCamera.cx: %.9g\nCamera.cy: %.9g\nCamera.bf: %.9g\n"
                % (Kf[0, 0], Kf[1, 1], Kf[0, 2], Kf[1, 2], float(bf)))
    trk = L.ref_tracking_new(path.encode())
    L.ref_set_alias_frame.argtypes = [C.c_void_p]
    L.ref_set_alias_frame(L.ref_tracking_last(trk))      # see name_of in oracle/ref_harness.cc
    lm = RR.LocalMap.__new__(RR.LocalMap); lm.h = L.ref_tracking_localmap()
    K9 = Kf.reshape(9).copy()
    out = []
    for t, ((Lf, Rf), disp) in enumerate(zip(frames, disps)):
        RR.LOG["fundamental"].clear()
        Lf = np.ascontiguousarray(Lf, np.uint8); Rf = np.ascontiguousarray(Rf, np.uint8)
        h, w = Lf.shape[:2]
        ch = 1 if Lf.ndim == 2 else Lf.shape[2]
        d = np.ascontiguousarray(disp, np.float32)
        bx = np.ascontiguousarray(np.asarray(boxes_per_frame[t], np.int32).reshape(-1, 4))
        map_before = lm.list()
        # the last frame's map points as they stand before the call: the veto's bad flags are read from them afterwards
        prev_last = RR.Frame(handle=L.ref_frame_copy(L.ref_tracking_last(trk))) if t > 0 else None
        L.ref_tracking_track(trk, RR._p(Lf), RR._p(Rf), w, h, ch, RR._p(d), RR._p(K9), float(bf), 0.1 * t, RR._p(bx), len(bx))
        assert L.ref_tracking_frame_num() == t + 1
        cur = RR.Frame(handle=L.ref_tracking_current(trk)); cur.shape = (h, w)
        last = RR.Frame(handle=L.ref_tracking_last(trk)); last.shape = (h, w)
        rec = dict(cur=cur.state(), map_before=map_before, last_after_match=None if prev_last is None else prev_last.state(),
                   F=dict(RR.LOG["fundamental"][-1])["F"] if RR.LOG["fundamental"] else None, last=last.state(), map=lm.list())
        rec["created"] = int((rec["map"]["create_id"] == t).sum())
        rec["erased"] = int((map_before["create_id"] <= t - 4).sum()) if t >= 4 else 0
        out.append(rec)
    L.ref_set_alias_frame(None)
    L.ref_tracking_free(trk)
    return out
