"""Pin the oracle against the live OpenCV in this container (skipped where cv2 is absent)."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


def cv_orb(img, nf):
    cv2.setUseOptimized(False)
    kp, desc = cv2.ORB_create(nfeatures=nf, scaleFactor=1.2, nlevels=8).detectAndCompute(img, None)
    return np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=O.KP_DTYPE), desc


@pytest.mark.parametrize("shape,seed,nf", [(synth.K_SHAPE, 2, 2000), (synth.K_SHAPE, 3, 500), (synth.K_SHAPE, 4, 4000),
                                           ((203, 317), 5, 300), (synth.H_SHAPE, 6, 8000)])
def test_orb_equals_cv2(shape, seed, nf):
    img = synth.texture(shape, seed)
    ref, rdesc = cv_orb(img, nf)
    kp, desc, _ = O.orb(img, nf)
    assert len(kp) == len(ref)
    for f in ("x", "y", "size", "angle", "response"):
        assert (kp[f].view(np.uint32) == ref[f].view(np.uint32)).all(), f
    assert (kp["octave"] == ref["octave"]).all()
    assert (desc == rdesc).all()


def test_pyramid_chain_equals_cv2_resize():
    img = synth.texture(synth.K_SHAPE, 9)
    lw, lh, _, _ = O.geometry(img.shape[1], img.shape[0], 8, 1.2, 2000)
    a = b = img
    for l in range(1, 8):
        a = cv2.resize(a, (int(lw[l]), int(lh[l])), interpolation=cv2.INTER_LINEAR_EXACT)
        b = O.resize(b, int(lw[l]), int(lh[l]))
        assert (a == b).all(), l


def test_sequence_frames_equal_cv2():
    seq = synth.Sequence(seed=0)
    for t in (0, 17):
        L, R = seq.frame(t)
        for img in (L, R):
            ref, rdesc = cv_orb(img, 2000)
            kp, desc, _ = O.orb(img, 2000)
            assert len(kp) == len(ref) and (desc == rdesc).all()
            assert (kp["x"] == ref["x"]).all() and (kp["angle"].view(np.uint32) == ref["angle"].view(np.uint32)).all()


def test_bgr2gray_equals_cvtcolor_on_every_colour():
    v = np.arange(256, dtype=np.uint8)
    for b in range(0, 256):
        img = np.empty((256, 256, 3), np.uint8)
        img[..., 0] = b; img[..., 1] = v[:, None]; img[..., 2] = v[None, :]
        assert (O.bgr2gray(img) == cv2.cvtColor(img, cv2.COLOR_BGR2GRAY)).all(), b


def test_orb_on_colour_input_equals_orb_on_converted_gray():
    """cv::ORB converts colour input itself: detectAndCompute(BGR) == oracle ORB on the oracle's gray."""
    col = synth.colourise(synth.texture((240, 400), 33), 1)
    ref, rdesc = cv_orb(col, 500)
    kp, desc, _ = O.orb(O.bgr2gray(col), 500)
    assert len(kp) == len(ref) and (desc == rdesc).all()
    for f in ("x", "y", "angle", "response"):
        assert (kp[f].view(np.uint32) == ref[f].view(np.uint32)).all(), f


@pytest.mark.parametrize("nl,sf,nf,shape", [(5, 1.3, 800, (376, 1241)), (3, 1.5, 300, (240, 400)), (8, 1.1, 1500, (376, 1241)),
                                            (1, 1.2, 200, (240, 400)), (6, 2.0, 400, (480, 640))])
def test_orb_equals_cv2_at_other_level_counts_and_scale_factors(nl, sf, nf, shape):
    """The oracle follows cv::ORB for ORBextractor.nLevels / scaleFactor other than KITTI's 8 / 1.2 as well (the cases
    tests/test_gpu_parity2.py runs on the device)."""
    img = synth.texture(shape, 17 + nl)
    cv2.setUseOptimized(False)
    kp, rdesc = cv2.ORB_create(nfeatures=nf, scaleFactor=sf, nlevels=nl).detectAndCompute(img, None)
    ref = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=O.KP_DTYPE)
    okp, desc, _ = O.orb(img, nf, scale=sf, nlevels=nl)
    assert len(okp) == len(ref) and (desc == rdesc).all() and (okp["octave"] == ref["octave"]).all()
    for f in ("x", "y", "size", "angle", "response"):
        assert (okp[f].view(np.uint32) == ref[f].view(np.uint32)).all(), f


# widths probed in cv2 4.13.0 (tools/probe_cv2_level_sizes.py): (image width, level, width of that level) at scale 1.2, all
# of them sizes where cols / scale sits within a float ulp of k + 0.5 and the plausible roundings disagree
CV2_LEVEL_WIDTHS = [
    (93, 1, 78), (117, 1, 98), (129, 1, 108), (189, 1, 158), (249, 1, 208), (285, 1, 238), (309, 1, 258), (381, 1, 318),
    (477, 1, 398), (573, 1, 478), (669, 1, 558), (765, 1, 638), (861, 1, 718), (957, 1, 798), (1053, 1, 878), (1149, 1, 958),
    (1245, 1, 1038), (1341, 1, 1118), (126, 2, 88), (198, 2, 138), (270, 2, 188), (342, 2, 237), (414, 2, 288), (486, 2, 338),
    (558, 2, 388), (630, 2, 437), (702, 2, 487), (774, 2, 538), (846, 2, 588), (918, 2, 638), (990, 2, 688), (1062, 2, 738),
    (324, 3, 187), (540, 3, 312), (756, 3, 437), (972, 3, 562), (1188, 3, 687), (1404, 3, 812)]


def test_level_sizes_follow_cv2_where_the_quotient_sits_on_a_half():
    """cvRound(cols / scale) as the pinned OpenCV build evaluates it — (float)cols * (1.f / scale) — not the quotient:
    249 / 1.2 -> 208 (the quotient rounds to 207).  The true quotient gets 27 of these 38 probes right."""
    for w, l, want in CV2_LEVEL_WIDTHS:
        lw, lh, _, _ = O.geometry(w, w, l + 1, 1.2, 500)
        assert lw[l] == want and lh[l] == want, (w, l, want, lw[l])
    # KITTI's shapes and the bench shapes are not among the affected sizes: both roundings agree on every level
    for n in (1241, 376, 1242, 375, 1226, 370, 2560, 720, 400, 240):
        lw, _, ls, _ = O.geometry(n, n, 8, 1.2, 500)
        assert [int(np.rint(np.float32(n) / s)) for s in ls] == list(lw), n


@pytest.mark.parametrize("shape,nf,nl", [((181, 249), 200, 6), ((297, 465), 600, 8), ((179, 558), 1000, 6)])
def test_orb_equals_cv2_at_sizes_on_a_half(shape, nf, nl):
    """Images whose level sizes depend on that rounding (found by tools/fuzz_orb_cv2.py): every field bit-equal."""
    img = synth.texture(shape, 77)
    cv2.setUseOptimized(False)
    kp, rdesc = cv2.ORB_create(nfeatures=nf, scaleFactor=1.2, nlevels=nl).detectAndCompute(img, None)
    ref = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=O.KP_DTYPE)
    okp, desc, _ = O.orb(img, nf, nlevels=nl)
    assert len(okp) == len(ref) > 0 and (desc == rdesc).all() and (okp["octave"] == ref["octave"]).all()
    for f in ("x", "y", "size", "angle", "response"):
        assert (okp[f].view(np.uint32) == ref[f].view(np.uint32)).all(), f


def test_a_slice_of_the_extraction_fuzz():
    """tools/fuzz_orb_cv2.py (random sizes, feature counts, level counts, scale factors, image kinds): 2000 seeds were run
    when this was written — 8 of the first 600 diverged, all through the level-size rounding above, none since; 16 of them here, the eight
    among them."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_orb_cv2 as Z
    total = 0
    for seed in (0, 1, 2, 3, 4, 5, 6, 7, 174, 182, 187, 274, 333, 467, 489, 540):
        msg, n = Z.run(seed)
        assert msg is None, msg
        total += n
    assert total > 5000


@pytest.mark.parametrize("sw,dw", [(3993, 3328), (1971, 1792), (1037, 768), (1535, 768), (1151, 768), (961, 768)])
def test_resize_coefficient_ties_follow_cv2(sw, dw):
    """INTER_LINEAR_EXACT coefficients that are exact ties ((fv - iv) * 256 = k + 0.5: v2(dst) - v2(src) = 8) depend on how
    the scale is formed; OpenCV divides one by dst / src.  3993 -> 3328 is the only such transition in the 1.2 pyramids of
    image dimensions up to 4095; the others are the ones of scale factors 1.1, 1.35, 2.0, 1.5 and 1.25."""
    rng = np.random.default_rng(sw)
    img = rng.integers(0, 256, (40, sw), dtype=np.uint8)
    assert (cv2.resize(img, (dw, 40), interpolation=cv2.INTER_LINEAR_EXACT) == O.resize(img, dw, 40)).all()
    assert (cv2.resize(img.T.copy(), (40, dw), interpolation=cv2.INTER_LINEAR_EXACT) == O.resize(img.T.copy(), 40, dw)).all()
