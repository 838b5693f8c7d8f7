"""Hand-made feature sets for the adversarial pins of the matchers (tests/test_ref_pin_adversarial.py,
tests/golden/make_golden_ref_adversarial.py): long chains of identical descriptors, exact ties, distances on the
thresholds of pass 1 (14 / 15) and pass 2 (29 / 30, second / best = 2.0 / 2.05), best = 0, keypoints off their epipolar
lines.  Pure numpy; test infrastructure."""
import numpy as np

SHAPE = (376, 1241)
CASES = {0: [], 1: [], 2: [[0, 1241, 0, 376]], 3: [[200, 700, 60, 250]]}        # seed -> offline YOLO boxes


def flip(d, bits):
    """d with the given bit positions inverted."""
    out = d.copy()
    for b in bits:
        out[b >> 3] ^= np.uint8(1 << (b & 7))
    return out


def at_distance(rng, d, k):
    return flip(d, rng.choice(256, k, replace=False))


def keypoints(rng, n):
    """n keypoints on a grid (7 floats each: x, y, size, angle, response, octave, class_id)."""
    k = np.zeros((n, 7), np.float32)
    cols = 50
    k[:, 0] = 30 + (np.arange(n) % cols) * 23.5
    k[:, 1] = 20 + (np.arange(n) // cols) * 31.25
    k[:, 2] = 31; k[:, 3] = rng.random(n) * 360; k[:, 4] = rng.random(n); k[:, 6] = -1
    return k


def build_case(seed):
    """(last frame's descriptors, current frame's descriptors) with the structures listed in the module docstring."""
    rng = np.random.default_rng(seed)
    n0 = n1 = 500        # frame::N is hard-wired (src/frame.cc:54) and createmappoint indexes keypoints_l up to it
    last = rng.integers(0, 256, (n0, 32), dtype=np.uint8)
    cur = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    X = last[0].copy()
    last[0:20] = X                                            # A: twenty identical rows ...
    slots = rng.permutation(n1)
    s = iter(slots)
    for _ in range(12):
        cur[next(s)] = X                                      # ... twelve identical columns (best = 0: claimed in index order)
    for _ in range(5):
        cur[next(s)] = at_distance(rng, X, 3)
    for _ in range(5):
        cur[next(s)] = at_distance(rng, X, 14)                # the last rows of the chain end on 14 (< 15) ...
    cur[next(s)] = at_distance(rng, X, 15)                    # ... and 15 is never claimed by pass 1
    for r in range(20, 40):                                   # B: exact ties — two columns at the same distance
        d = int(rng.integers(1, 14))
        cur[next(s)] = at_distance(rng, last[r], d)
        cur[next(s)] = at_distance(rng, last[r], d)
    for r in range(40, 60):                                   # C: thresholds of pass 1
        cur[next(s)] = at_distance(rng, last[r], 14 if r % 2 else 15)
    for r in range(60, 100):                                  # D: thresholds of pass 2 (rows pass 1 leaves alone)
        best = (29, 30, 20, 20)[r % 4]
        second = (200, 200, 40, 41)[r % 4]                    # 40 / 20 = 2.0 is not > 2; 41 / 20 is
        base = at_distance(rng, last[r], best)
        cur[next(s)] = base
        if second < 100:
            # a second column at exactly `second` from the row: flip bits that base did not touch
            same = np.nonzero(np.unpackbits(base ^ last[r], bitorder="little") == 0)[0]
            cur[next(s)] = flip(last[r], rng.choice(same, second, replace=False))
    for r in range(100, 200):                                 # E: ordinary near-duplicates
        cur[next(s)] = last[r] ^ np.packbits(rng.random(256) < 0.02, bitorder="little")
    return last, cur


def feature_sets(seed):
    """((keypoints, descriptors) of the last frame, (keypoints, descriptors) of the current frame)."""
    rng = np.random.default_rng(100 + seed)
    d_last, d_cur = build_case(seed)
    k_last = keypoints(rng, len(d_last))
    k_cur = keypoints(rng, len(d_cur))
    k_cur[:, 0] -= 3.0                                        # a small horizontal motion ...
    off = rng.random(len(d_cur)) < 0.3
    k_cur[off, 1] += rng.uniform(-4, 4, off.sum()).astype(np.float32)   # ... and rows off their epipolar lines
    return (k_last, d_last), (k_cur, d_cur)


def run_reference(seed, boxes, K, bf):
    """The two hand-made frames through the reference's own code (oracle/ref.py:run_two_frames): its cv::ORB calls are
    answered with the hand-made sets, the dense disparity is 10 px everywhere (every keypoint gets a map point)."""
    from oracle import ref as R
    feats = dict(zip((10, 20), feature_sets(seed)))
    imgs = {}
    for tag in feats:
        im = np.full(SHAPE, 128, np.uint8); im[0, 0] = tag    # the tag tells the hook which frame asks
        imgs[tag] = im

    def orb(img, what):
        return feats[int(img.reshape(img.shape[0], -1)[0, 0])]

    disp = np.full(SHAPE, 10, np.float32)
    R.ORB_OVERRIDE = orb
    try:
        return R.run_two_frames(((imgs[10], imgs[10]), (imgs[20], imgs[20])), (disp, disp), K, bf, boxes)
    finally:
        R.ORB_OVERRIDE = None


GOLDEN_KEYS = {"before.last": ("N", "mp_create_id", "desc"), "before.map": ("create_id", "idx", "desc"),
               "cur": ("N", "kps", "desc", "match_score", "mp_create_id", "mp_idx"), "last": ("kps", "mp_bad")}


def pack_run(run, prefix):
    """What check_run (tests/test_ref_pin.py) reads of a run, flattened for an .npz."""
    out = {prefix + "F": np.zeros((0, 0)) if run["F"]["F"] is None else run["F"]["F"], prefix + "created": np.int64(run["created"])}
    for path, keys in GOLDEN_KEYS.items():
        node = run
        for part in path.split("."):
            node = node[part]
        for k in keys:
            out[prefix + path + "." + k] = np.asarray(node[k])
    return out


def unpack_run(g, prefix):
    F = g[prefix + "F"]
    run = {"F": {"F": None if F.size == 0 else F}, "created": int(g[prefix + "created"]), "before": {}}
    for path, keys in GOLDEN_KEYS.items():
        node = {k: (int(g[prefix + path + "." + k]) if k == "N" else g[prefix + path + "." + k]) for k in keys}
        parts = path.split(".")
        if len(parts) == 2:
            run[parts[0]][parts[1]] = node
        else:
            run[path] = node
    return run


def sequence_sets(seed, n):
    """n frames of 500 hand-made features for the tracker pins: every frame re-observes most of the previous frame's
    descriptors (0-3 bits flipped, shuffled positions in the descriptor matrix), brings back descriptors of two and three
    frames ago with ~20 bits flipped (pass-2 material: best < 30, ratio > 2 against everything else), carries a block of
    identical descriptors (a claim chain in both passes), and fills up with fresh random rows."""
    rng = np.random.default_rng(1000 + seed)
    N = 500
    sets = []
    for t in range(n):
        d = rng.integers(0, 256, (N, 32), dtype=np.uint8)
        slots = iter(rng.permutation(N))
        if t > 0:
            prev = sets[t - 1][1]
            for r in rng.choice(N, 280, replace=False):
                d[next(slots)] = at_distance(rng, prev[r], int(rng.integers(0, 4)))
            for back in (2, 3):
                if t >= back:
                    old = sets[t - back][1]
                    for r in rng.choice(N, 40, replace=False):
                        d[next(slots)] = at_distance(rng, old[r], int(rng.integers(16, 29)))
        chain = (sets[0][1][7] if t else d[7]).copy()          # the same block value in every frame
        for _ in range(10):
            d[next(slots)] = chain
        k = keypoints(rng, N)
        k[:, 0] += rng.uniform(-2, 2, N).astype(np.float32) - 2.0 * t
        k[:, 1] += rng.uniform(-1.5, 1.5, N).astype(np.float32)
        sets.append((k, d))
    return sets


def run_reference_sequence(seed, n, boxes_of, K, bf, runner=None, **kw):
    """The hand-made sequence through the reference's frame loop (oracle/ref.py:run_sequence by default; pass
    oracle.ref_g2o.run_tracking + tmpdir for Tracking::Track itself)."""
    from oracle import ref as R
    sets = sequence_sets(seed, n)
    imgs = []
    for t in range(n):
        im = np.full(SHAPE, 128, np.uint8); im[0, 0] = 10 + t
        imgs.append(im)

    def orb(img, what):
        return sets[int(img.reshape(img.shape[0], -1)[0, 0]) - 10]

    disp = np.full(SHAPE, 10, np.float32)
    mods = [R]
    if runner is not None:
        from oracle import ref_g2o as RG
        mods.append(RG._ref_module())                          # the module instance bound to libsvo_ref_g2o.so
    for m in mods:
        m.ORB_OVERRIDE = orb
    try:
        run = runner or R.run_sequence
        return run([(im, im) for im in imgs], [disp] * n, K, bf, [boxes_of(t) for t in range(n)], **kw)
    finally:
        for m in mods:
            m.ORB_OVERRIDE = None
