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
