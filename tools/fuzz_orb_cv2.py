"""Fuzz of the extraction pin (test infrastructure): random image sizes, feature counts, level counts / scale factors and
image kinds (the bench texture, uniform noise, smooth gradients with sparse dots, a checkerboard, low-contrast texture,
real street images when /root/reference is present) through cv2.ORB_create(...).detectAndCompute and through the oracle;
every field must be bit-equal.  Prints the cases that diverge.

    python tools/fuzz_orb_cv2.py [first_seed] [count]
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200")]
import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

REAL = "/root/reference/Thirdparty/libelas/img"


def image(rng, kind, h, w):
    if kind == 0:
        return synth.texture((h, w), int(rng.integers(1 << 30)))
    if kind == 1:
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == 2:
        g = (np.add.outer(np.arange(h) * rng.uniform(0, 0.4), np.arange(w) * rng.uniform(0, 0.2)) % 256).astype(np.uint8)
        n = int(rng.integers(0, 200))
        g[rng.integers(0, h, n), rng.integers(0, w, n)] = rng.integers(0, 256, n)
        return g
    if kind == 3:
        c = int(rng.integers(3, 24))
        return (((np.arange(h)[:, None] // c + np.arange(w)[None, :] // c) & 1) * int(rng.integers(20, 255))).astype(np.uint8)
    if kind == 4:
        t = synth.texture((h, w), int(rng.integers(1 << 30))).astype(np.int32)
        return np.clip(128 + (t - 128) // int(rng.integers(3, 9)), 0, 255).astype(np.uint8)
    files = sorted(f for f in os.listdir(REAL) if f.endswith(".pgm")) if os.path.isdir(REAL) else []
    if not files:
        return synth.texture((h, w), int(rng.integers(1 << 30)))
    im = cv2.imread(os.path.join(REAL, files[int(rng.integers(len(files)))]), cv2.IMREAD_GRAYSCALE)
    y0 = int(rng.integers(0, max(1, im.shape[0] - h))); x0 = int(rng.integers(0, max(1, im.shape[1] - w)))
    return np.ascontiguousarray(im[y0:y0 + h, x0:x0 + w])


def run(seed):
    rng = np.random.default_rng(seed)
    kind = int(rng.integers(0, 6))
    h, w = int(rng.integers(70, 500)), int(rng.integers(70, 900))
    nl = int(rng.choice([8, 8, 8, 1, 2, 4, 6])); sf = float(rng.choice([1.2, 1.2, 1.2, 1.1, 1.35, 1.5, 2.0]))
    nf = int(rng.choice([50, 200, 500, 1000, 2000, 4000]))
    while min(h, w) / sf ** (nl - 1) < 70:       # every level must hold the 31-px border twice (cv::ORB's own limit)
        nl -= 1
    img = image(rng, kind, h, w)
    h, w = img.shape
    cv2.setUseOptimized(False)
    kp, rdesc = cv2.ORB_create(nfeatures=nf, scaleFactor=sf, nlevels=nl).detectAndCompute(img, None)
    ref = np.array([(p.pt[0], p.pt[1], p.size, p.angle, p.response, p.octave) for p in kp], dtype=O.KP_DTYPE)
    okp, desc, _ = O.orb(img, nf, scale=sf, nlevels=nl)
    tag = "seed %d kind %d %dx%d nf %d levels %d scale %.2f: %d keypoints" % (seed, kind, w, h, nf, nl, sf, len(ref))
    if len(okp) != len(ref):
        return tag + " COUNT %d" % len(okp), 0
    if len(ref) == 0:
        return None, 0
    for f in ("x", "y", "size", "angle", "response"):
        if not (okp[f].view(np.uint32) == ref[f].view(np.uint32)).all():
            return tag + " FIELD " + f, 0
    if not (okp["octave"] == ref["octave"]).all() or not (desc == rdesc).all():
        return tag + " OCTAVE/DESC", 0
    return None, len(ref)


if __name__ == "__main__":
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    bad, total = [], 0
    for s in range(first, first + count):
        msg, n = run(s)
        total += n
        if msg:
            bad.append(msg); print("DIVERGES:", msg, file=sys.stderr)
    print("seeds %d..%d: %d divergent; %d keypoints compared bit for bit" % (first, first + count - 1, len(bad), total))
