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
