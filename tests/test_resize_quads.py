"""CPU check of the quad-table resize (pyramid.cu:k_resize_q, csrc/resize_quads.h).

The kernel cannot run here, so its arithmetic is replayed on the CPU by tests/host_models/resize_quads_model.cc: the SAME
host table code, the kernel's thread/strip/register-role structure, and plain-C++ restatements of the three intrinsics
it uses.  Expectation: cv2.resize(INTER_LINEAR_EXACT), which is what cv::ORB's pyramid calls.  The
GPU parity tests (tests/test_gpu_parity.py, pyramid taps) check the kernel itself against the oracle.
"""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def model(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("rq") / "librq_model.so")
    subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-o", so,
                           os.path.join(HERE, "host_models", "resize_quads_model.cc")])
    L = C.CDLL(so)
    L.resize_model.restype = C.c_int
    L.resize_model.argtypes = [C.c_void_p] + [C.c_int] * 6 + [C.c_void_p, C.c_int]
    return L


def _level_sizes(w, h, nlevels, sf):
    out = []
    for l in range(nlevels):
        s = np.float32(np.float64(np.float32(sf)) ** l)
        inv = np.float32(1) / s                          # cvRound(cols * (1 / scale)), as orb_geometry (csrc/svo_api.cu)
        out.append((int(np.rint(np.float32(w) * inv)), int(np.rint(np.float32(h) * inv))))
    return out


def _run(model, src_img, dw, dh, rs):
    sh, sw = src_img.shape
    spitch = (sw + 15) // 16 * 16
    dpitch = (dw + 15) // 16 * 16
    src = np.full((sh + 1, spitch), 0xA5, np.uint8)     # padding holds junk: it must never reach a result
    src[:sh, :sw] = src_img
    dst = np.zeros((dh, dpitch), np.uint8)
    r = model.resize_model(src.ctypes.data, sw, sh, spitch, dw, dh, dpitch, dst.ctypes.data, rs)
    return r, dst


@pytest.mark.parametrize("w,h,nlevels,sf", [(1241, 376, 8, 1.2), (2560, 720, 8, 1.2), (1242, 375, 8, 1.2), (1226, 370, 8, 1.2),
                                            (400, 240, 8, 1.2), (317, 203, 6, 1.2), (640, 480, 5, 1.5), (640, 480, 4, 1.1),
                                            (800, 600, 3, 1.9), (4095, 64, 3, 1.2),
                                            (249, 181, 6, 1.2), (465, 297, 8, 1.2), (558, 341, 8, 1.2)])   # level sizes on a half
def test_quad_table_equals_cv2(model, w, h, nlevels, sf):
    rng = np.random.default_rng(w * 31 + h)
    sizes = _level_sizes(w, h, nlevels, sf)
    img = rng.integers(0, 256, (h, w), dtype=np.uint8)
    for l in range(1, nlevels):
        dw, dh = sizes[l]
        for rs in (8, 4):
            r, dst = _run(model, img, dw, dh, rs)
            assert r == 0, (l, rs, r)                      # quad table applies and equals the per-byte form, padding 0
        ref = cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(dst[:, :dw], ref), l
        assert not dst[:, dw:].any()
        img = ref


def test_steep_scale_falls_back(model):
    """Scale factors whose taps do not fit the 8-byte window make build_resize_quads report it (the per-byte kernel
    then runs for that level); factors up to cv::ORB's practical range never do."""
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (300, 900), dtype=np.uint8)
    r, _ = _run(model, img, 300, 100, 8)                   # scale 3.0
    assert r == -1
    r, dst = _run(model, img, 451, 151, 8)                 # just under 2.0
    assert r == 0
    assert np.array_equal(dst[:, :451], cv2.resize(img, (451, 151), interpolation=cv2.INTER_LINEAR_EXACT))


def test_extremes(model):
    """Saturated images (largest sums: every result byte must be exact, no carry into a neighbour), constant rows."""
    for val in (0, 255):
        img = np.full((376, 1241), val, np.uint8)
        r, dst = _run(model, img, 1034, 313, 8)
        assert r == 0 and (dst[:, :1034] == val).all()
