"""Fuzz of the matcher pins (test infrastructure; needs /root/reference or a prebuilt oracle/_ref): random hand-made
feature sets — duplicate blocks of random sizes, distances drawn around the thresholds, random boxes, random keypoint
positions — through the reference's own poseEstimationPnP, each run compared with the oracle by the checks of
tests/test_ref_pin.py.  Prints the seeds that diverge (none so far).

    python tools/fuzz_ref_pin.py [first_seed] [count]
    python tools/fuzz_ref_pin.py --sequences [first_seed] [count]     # hand-made 8-frame sequences with random boxes through
                                                                      # the reference's frame loop against oracle/track.py
"""
import contextlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200"), os.path.join(ROOT, "tests")]
import adversarial_sets as A  # noqa: E402
import synth  # noqa: E402
from oracle import ref as R  # noqa: E402
from test_ref_pin import check_run  # noqa: E402

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
BF = np.float32(CAL["bf"])


@contextlib.contextmanager
def quiet():
    """The reference prints every epipolar distance it measures (std::cout): silence file descriptor 1 meanwhile."""
    sys.stdout.flush()
    keep, null = os.dup(1), os.open(os.devnull, os.O_WRONLY)
    os.dup2(null, 1)
    try:
        yield
    finally:
        os.dup2(keep, 1); os.close(keep); os.close(null)


def random_case(seed):
    rng = np.random.default_rng(seed)
    N = 500
    last = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    cur = rng.integers(0, 256, (N, 32), dtype=np.uint8)
    for _ in range(int(rng.integers(1, 5))):                      # blocks of identical rows
        a, n = int(rng.integers(0, N - 40)), int(rng.integers(2, 40))
        last[a:a + n] = last[a]
    slots = iter(rng.permutation(N))
    for r in rng.choice(N, 330, replace=False):
        kind = rng.integers(0, 6)
        d = int((0, rng.integers(0, 5), rng.integers(12, 18), rng.integers(26, 33), rng.integers(0, 30), rng.integers(0, 60))[kind])
        cur[next(slots)] = A.at_distance(rng, last[r], d)
        if rng.random() < 0.3:                                    # a rival column for the ratio test / a tie
            try:
                cur[next(slots)] = A.at_distance(rng, last[r], int(d * rng.choice([1.0, 2.0, 2.05, 3.0])))
            except StopIteration:
                break
    kl, kc = A.keypoints(rng, N), A.keypoints(rng, N)
    for k in (kl, kc):
        k[:, 0] = rng.uniform(5, 1236, N); k[:, 1] = rng.uniform(5, 371, N)
    kc[:, :2] = kl[rng.permutation(N), :2] + rng.normal(0, 1.5, (N, 2))
    kc[:, 0] = np.clip(kc[:, 0], 1, 1239); kc[:, 1] = np.clip(kc[:, 1], 1, 374)
    boxes = []
    for _ in range(int(rng.integers(0, 4))):
        x0, y0 = int(rng.integers(0, 1000)), int(rng.integers(0, 250))
        boxes.append([x0, x0 + int(rng.integers(50, 600)), y0, y0 + int(rng.integers(40, 300))])
    return (kl.astype(np.float32), last), (kc.astype(np.float32), cur), boxes


def run(seed):
    f_last, f_cur, boxes = random_case(seed)
    feats = {10: f_last, 20: f_cur}
    imgs = {t: np.full(A.SHAPE, 128, np.uint8) for t in feats}
    for t in feats:
        imgs[t][0, 0] = t
    R.ORB_OVERRIDE = lambda img, what: feats[int(img.reshape(img.shape[0], -1)[0, 0])]
    disp = np.full(A.SHAPE, 10, np.float32)
    disp[:, ::7] = -1                                             # some keypoints without depth: no map point
    try:
        with quiet():
            out = R.run_two_frames(((imgs[10], imgs[10]), (imgs[20], imgs[20])), (disp, disp), K, BF, boxes)
    finally:
        R.ORB_OVERRIDE = None
    if out["F"]["F"] is None:
        out["F"]["F"] = np.zeros((3, 3))
    p1, p2 = check_run(out, boxes)
    return int(p1["row_claimed"].sum()), int(p1["row_bad"].sum()), int(p2["row_claimed"].sum())


def run_sequence(seed, n=8):
    """tests/adversarial_sets.py:sequence_sets with random boxes per frame; returns (vetoes, pass-1 claims, pass-2 claims)."""
    from test_oracle_track import check, replay
    rng = np.random.default_rng(seed)
    bx = []
    for _ in range(n):
        b = []
        for _ in range(int(rng.integers(0, 3))):
            x0, y0 = int(rng.integers(0, 1000)), int(rng.integers(0, 250))
            b.append([x0, x0 + int(rng.integers(50, 500)), y0, y0 + int(rng.integers(40, 250))])
        bx.append(b)
    boxes_of = lambda t: bx[t]
    with quiet():
        recs = A.run_reference_sequence(seed, n, boxes_of, K, BF)
    return check(recs, replay(recs, boxes_of))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--sequences":
        first = int(sys.argv[2]) if len(sys.argv) > 2 else 0
        count = int(sys.argv[3]) if len(sys.argv) > 3 else 20
        bad, tot = [], np.zeros(3, np.int64)
        for s in range(first, first + count):
            try:
                tot += run_sequence(s)
            except AssertionError as e:
                bad.append(s); print("seed", s, "DIVERGES:", str(e)[:300], file=sys.stderr)
        print("sequences %d..%d: %d divergent %s; vetoed %d, pass-1 claims %d, pass-2 claims %d" % (first, first + count - 1, len(bad), bad, *tot))
        sys.exit(0)
    first = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    count = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    bad = []
    tot = np.zeros(3, np.int64)
    for s in range(first, first + count):
        try:
            tot += run(s)
        except AssertionError as e:
            bad.append(s); print("seed", s, "DIVERGES:", e, file=sys.stderr)
    print("seeds %d..%d: %d divergent %s; pass-1 claims %d, vetoed %d, pass-2 claims %d" % (first, first + count - 1, len(bad), bad, *tot))
