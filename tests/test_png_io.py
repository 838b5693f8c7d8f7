"""Input staging: svo_png_decode against cv2.imread(..., IMREAD_UNCHANGED), which is how
main.cpp:160-162 reads the KITTI frames.  Host code: runs without a GPU."""
import numpy as np
import pytest

import svo
import synth

cv2 = pytest.importorskip("cv2")


def roundtrip(img, params=()):
    ok, enc = cv2.imencode(".png", img, list(params))
    assert ok
    data = enc.tobytes()
    want = cv2.imdecode(np.frombuffer(data, np.uint8), cv2.IMREAD_UNCHANGED)
    got = svo.png_decode(data)
    assert got.dtype == want.dtype and got.shape == want.shape and (got == want).all()
    return data


@pytest.mark.parametrize("strategy", [cv2.IMWRITE_PNG_STRATEGY_DEFAULT, cv2.IMWRITE_PNG_STRATEGY_FILTERED, cv2.IMWRITE_PNG_STRATEGY_RLE,
                                      cv2.IMWRITE_PNG_STRATEGY_HUFFMAN_ONLY, cv2.IMWRITE_PNG_STRATEGY_FIXED])
def test_gray_kitti_frame(strategy):
    L, R = synth.Sequence(synth.K_SHAPE, seed=2).frame(3)
    for level in (1, 6):
        roundtrip(L, (cv2.IMWRITE_PNG_STRATEGY, strategy, cv2.IMWRITE_PNG_COMPRESSION, level))


def test_colour_alpha_16bit_and_odd_sizes():
    rng = np.random.default_rng(0)
    base = synth.texture((203, 317), 5)
    bgr = np.stack([base, np.roll(base, 7, 0), np.roll(base, 11, 1)], 2)
    roundtrip(bgr)                                                                  # RGB in the file, BGR in memory
    roundtrip(np.dstack([bgr, rng.integers(0, 256, base.shape, dtype=np.uint8)]))   # RGBA -> BGRA
    roundtrip((base.astype(np.uint16) * 257 + rng.integers(0, 255, base.shape)).astype(np.uint16))   # 16-bit gray (depth PNGs)
    roundtrip(rng.integers(0, 256, (1, 1), dtype=np.uint8))
    roundtrip(rng.integers(0, 256, (7, 3, 3), dtype=np.uint8))
    smooth = (np.add.outer(np.arange(240), np.arange(400)) // 3).astype(np.uint8)  # smooth ramps make the encoder pick every filter
    roundtrip(smooth)
    roundtrip(np.stack([smooth, smooth.T[:240, :240].repeat(2, 1)[:, :400], 255 - smooth], 2))


def test_decodes_into_a_strided_destination_and_rejects_what_it_cannot_read():
    img = synth.texture((60, 90), 1)
    data = roundtrip(img)
    dst = np.zeros((60, 128), np.uint8)
    out = svo.png_decode(data, out=dst[:, :90])
    assert (out == img).all() and (dst[:, 90:] == 0).all()
    with pytest.raises(svo.SvoError):
        svo.png_decode(b"not a png at all, just some bytes that are long enough to pass the size check")
    with pytest.raises(svo.SvoError):
        svo.png_decode(data[:len(data) // 2])                                       # truncated stream
    ok, enc = cv2.imencode(".png", (img > 128).astype(np.uint8) * 255, [cv2.IMWRITE_PNG_BILEVEL, 1])
    with pytest.raises(svo.SvoError):
        svo.png_decode(enc.tobytes())                                               # 1-bit depth
