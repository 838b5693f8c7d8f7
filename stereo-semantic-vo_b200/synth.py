"""Seeded synthetic stereo inputs.  numpy only.

These are INPUT generators for tests and bench.py; nothing here is on the
product path.  Images are corner-rich so every pyramid level saturates its
retainBest quota, which is what exercises the order-sensitive selection.
"""
import numpy as np

K_SHAPE = (376, 1241)   # KITTI
H_SHAPE = (720, 2560)   # high-res config

# Stereo/KITTI04-12.yaml:8-11,25 and Stereo/KITTI00-02.yaml:8-11,25
KITTI_04_12 = dict(fx=707.0912, fy=707.0912, cx=601.8873, cy=183.1104, bf=379.8145)
KITTI_00_02 = dict(fx=718.856, fy=718.856, cx=607.1928, cy=185.2157, bf=386.1448)


def _blur3(a):
    """3x3 Gaussian sigma 0.8, float32, edge-replicated."""
    k = np.exp(-0.5 * (np.arange(-1, 2) / 0.8) ** 2); k = (k / k.sum()).astype(np.float32)
    p = np.pad(a, 1, mode="edge")
    h = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    return k[0] * h[:-2] + k[1] * h[1:-1] + k[2] * h[2:]


def texture(shape, seed, rect_density=1500.0 / (376 * 1241)):
    """gray 128 + random rectangles + N(0,3) noise, 3x3 blur, clipped to u8."""
    rng = np.random.default_rng(seed)
    h, w = shape
    img = np.full((h, w), 128.0, np.float32)
    n = int(round(rect_density * h * w))
    xs = rng.integers(0, w, n); ys = rng.integers(0, h, n)
    ws = rng.integers(4, 60, n); hs = rng.integers(4, 60, n)
    ds = rng.integers(-90, 90, n)
    for x, y, rw, rh, d in zip(xs, ys, ws, hs, ds):
        img[y:y + rh, x:x + rw] += d
    img += rng.normal(0, 3, img.shape).astype(np.float32)
    return np.clip(np.rint(_blur3(img)), 0, 255).astype(np.uint8)


def stereo_pair(shape=K_SHAPE, seed=0, noise=2.0):
    """Left texture and a right view warped by a piece-wise constant disparity in [2, 48] px."""
    rng = np.random.default_rng(seed + 7919)
    h, w = shape
    left = texture(shape, seed)
    # piece-wise constant disparity on a coarse grid of blocks
    bh, bw = 47, 73
    gd = rng.uniform(2.0, 48.0, ((h + bh - 1) // bh, (w + bw - 1) // bw)).astype(np.float32)
    disp = np.kron(gd, np.ones((bh, bw), np.float32))[:h, :w]
    # right(x) = left(x + d): a point at uL appears at uR = uL - d
    xr = np.arange(w, dtype=np.float32)[None, :] + disp
    x0 = np.clip(np.floor(xr).astype(np.int64), 0, w - 1); x1 = np.clip(x0 + 1, 0, w - 1)
    t = (xr - np.floor(xr)).astype(np.float32)
    lf = left.astype(np.float32)
    rows = np.arange(h)[:, None]
    right = (1 - t) * lf[rows, x0] + t * lf[rows, x1]
    right += rng.normal(0, noise, right.shape).astype(np.float32)
    return left, np.clip(np.rint(right), 0, 255).astype(np.uint8), disp


class Sequence:
    """A moving camera over a large master texture: frame t is a similarity warp of the master."""

    def __init__(self, shape=K_SHAPE, seed=0, master_scale=2.0, noise=2.0):
        h, w = shape
        self.shape = shape
        self.seed = seed
        self.noise = noise
        self.master = texture((int(h * master_scale * 2), int(w * master_scale * 1.65)), seed + 101).astype(np.float32)
        rng = np.random.default_rng(seed + 3)
        mh, mw = self.master.shape
        bh, bw = 61, 97
        gd = rng.uniform(2.0, 48.0, ((mh + bh - 1) // bh, (mw + bw - 1) // bw)).astype(np.float32)
        self.disp_master = np.kron(gd, np.ones((bh, bw), np.float32))[:mh, :mw]

    def _sample(self, xs, ys):
        mh, mw = self.master.shape
        x0 = np.floor(xs).astype(np.int64); y0 = np.floor(ys).astype(np.int64)
        tx = (xs - x0).astype(np.float32); ty = (ys - y0).astype(np.float32)
        x0 %= mw; y0 %= mh
        x1 = (x0 + 1) % mw; y1 = (y0 + 1) % mh
        m = self.master
        return ((1 - ty) * ((1 - tx) * m[y0, x0] + tx * m[y0, x1]) + ty * ((1 - tx) * m[y1, x0] + tx * m[y1, x1]))

    def frame(self, t):
        """-> (left u8, right u8).  tx += 3 px per frame; zoom breathes in [1, 1.25]."""
        h, w = self.shape
        rng = np.random.default_rng(self.seed * 100003 + t)
        zoom = 1.0 + 0.125 * (1 - np.cos(2 * np.pi * t / 250.0))
        ox, oy = 3.0 * t, 0.7 * t
        u = np.arange(w, dtype=np.float32)[None, :]; v = np.arange(h, dtype=np.float32)[:, None]
        xs = ox + u / zoom + 0 * v; ys = oy + v / zoom + 0 * u
        left = self._sample(xs, ys)
        mh, mw = self.master.shape
        d = self.disp_master[np.floor(ys).astype(np.int64) % mh, np.floor(xs).astype(np.int64) % mw]
        right = self._sample(xs + d / zoom, ys)
        left = left + rng.normal(0, self.noise, left.shape).astype(np.float32)
        right = right + rng.normal(0, self.noise, right.shape).astype(np.float32)
        return (np.clip(np.rint(left), 0, 255).astype(np.uint8), np.clip(np.rint(right), 0, 255).astype(np.uint8))


def local_map(desc_frames, rows=5000, seed=0):
    """5k-row local map: descriptors of the previous frames, padded with bit-flipped copies
    (Bernoulli p in {0.02, 0.1, 0.5}) up to `rows`, explicit row order."""
    rng = np.random.default_rng(seed)
    base = np.concatenate([np.asarray(d, np.uint8).reshape(-1, 32) for d in desc_frames], 0)
    if len(base) >= rows:
        return np.ascontiguousarray(base[rng.permutation(len(base))[:rows]])
    out = [base]
    need = rows - len(base)
    ps = (0.02, 0.1, 0.5)
    k = 0
    while need > 0:
        take = min(need, len(base))
        src = base[rng.permutation(len(base))[:take]]
        flips = np.packbits(rng.random((take, 256)) < ps[k % 3], axis=1, bitorder="little")
        out.append(src ^ flips)
        need -= take; k += 1
    return np.ascontiguousarray(np.concatenate(out, 0))


def rodrigues(w):
    """Rotation matrix of the axis-angle vector w (float64)."""
    w = np.asarray(w, np.float64)
    th = np.linalg.norm(w)
    Kx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + Kx
    return np.eye(3) + np.sin(th) / th * Kx + (1 - np.cos(th)) / th ** 2 * (Kx @ Kx)


def pose_problem(n, seed, outlier_frac=0.3, noise=0.5, cal=None, shape=K_SHAPE, rot=0.05, trans=0.3):
    """Synthetic input of the pose stage (src/pnpmatch.cc:211-227): n map points seen by a camera at a
    random pose (Xc = R Xw + t), pixel noise N(0, noise) and a fraction of gross mismatches.
    -> (Xw[n,3] f32, obs[n,2] f32, K=(fx,fy,cx,cy), R, t, outlier flags)."""
    cal = cal or KITTI_04_12
    fx, fy, cx, cy = cal["fx"], cal["fy"], cal["cx"], cal["cy"]
    rng = np.random.default_rng(seed)
    R = rodrigues(rng.normal(0, rot, 3)); t = rng.normal(0, trans, 3)
    uv = np.stack([rng.uniform(0, shape[1], n), rng.uniform(0, shape[0], n)], 1)
    z = rng.uniform(4, 60, n)
    Xc = np.stack([(uv[:, 0] - cx) / fx * z, (uv[:, 1] - cy) / fy * z, z], 1)
    Xw = (Xc - t) @ R
    obs = uv + rng.normal(0, noise, (n, 2))
    bad = rng.random(n) < outlier_frac
    obs[bad] = np.stack([rng.uniform(0, shape[1], int(bad.sum())), rng.uniform(0, shape[0], int(bad.sum()))], 1)
    return Xw.astype(np.float32), obs.astype(np.float32), (fx, fy, cx, cy), R, t, bad


def colourise(gray, seed):
    """Interleaved BGR image whose channels are seeded per-pixel perturbations of `gray` (a colour KITTI frame
    stand-in: image_2 / image_3 read with cv::imread(.., UNCHANGED), main.cpp:160-161)."""
    rng = np.random.default_rng(seed)
    g = gray.astype(np.int16)
    out = np.empty(gray.shape + (3,), np.uint8)
    gains = (0.85, 1.0, 1.15)
    for c in range(3):
        out[..., c] = np.clip(g * gains[c] + rng.integers(-12, 13, gray.shape) + (8, 0, -8)[c], 0, 255).astype(np.uint8)
    return out


def dense_disparity(shape, seed, holes=0.15, zeros=0.05):
    """A dense CV_32F disparity image standing in for frame::MB's output (src/frame.cc:82-91): piece-wise constant
    values in [2, 48], with a fraction of -1 pixels ("no disparity", which makes computekeypoint_r's rx stick,
    src/frame.cc:133) and of exact zeros (which disp2Depth skips, src/frame.cc:158)."""
    h, w = shape
    r = np.random.default_rng(seed + 4241)
    g = r.uniform(2, 48, ((h + 46) // 47, (w + 72) // 73)).astype(np.float32)
    d = np.kron(g, np.ones((47, 73), np.float32))[:h, :w].copy()
    m = r.random((h, w))
    d[m < holes] = -1
    d[(m >= holes) & (m < holes + zeros)] = 0
    return d
