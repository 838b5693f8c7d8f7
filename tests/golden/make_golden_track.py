"""Records what the REFERENCE'S OWN CODE computes over a seven-frame sequence driven like Tracking::Track
(src/Tracking.cc:184-250; oracle/ref.py:run_sequence over oracle/_ref/libsvo_ref.so = src/pnpmatch.cc, src/frame.cc,
src/mappoint.cc compiled unmodified), as a golden fixture for tests/test_oracle_track.py — /root/reference does not
exist on the GPU box, where the fixture pins oracle/track.py, the oracle of the device-resident tracker state.

Per frame: what the matching stage saw (keypoints, descriptors, depth at the keypoints, boxes, F, the local map's own
scan order) and what came out (CurrentFrame->MapPoints names, match_score, bad flags, points created, the map).

    python tests/golden/make_golden_track.py        # writes tests/golden/track_seq5.npz
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200"), os.path.join(ROOT, "tests")]
import test_oracle_track as TT  # noqa: E402

SEED, NF = 5, 7


def boxes_of(t):
    return TT.BOXES if t % 2 == 1 else TT.BOXES[:1]


def main():
    recs = TT.reference_run(SEED, NF, boxes_of)
    d = dict(seed=SEED, n=NF)
    for t, r in enumerate(recs):
        cur, last = r["cur"], r["last"]
        d["kps_%d" % t] = cur["kps"][:, :2]; d["desc_%d" % t] = cur["desc"]; d["depth_%d" % t] = cur["depth_at_kp"]
        d["boxes_%d" % t] = np.asarray(boxes_of(t), np.int32).reshape(-1, 4)
        d["F_%d" % t] = np.zeros((0, 0)) if r["F"] is None else np.asarray(r["F"], np.float64)
        for k in ("create_id", "idx", "bad"):
            d["mapb_%s_%d" % (k, t)] = r["map_before"][k]; d["map_%s_%d" % (k, t)] = r["map"][k]
        d["cur_mp_create_%d" % t] = cur["mp_create_id"]; d["cur_mp_idx_%d" % t] = cur["mp_idx"]
        d["match_score_%d" % t] = cur["match_score"]
        d["last_mp_create_%d" % t] = last["mp_create_id"]; d["last_mp_idx_%d" % t] = last["mp_idx"]
        d["prev_bad_%d" % t] = np.zeros(0, np.uint8) if r["last_after_match"] is None else r["last_after_match"]["mp_bad"]
        d["created_%d" % t] = r["created"]; d["erased_%d" % t] = r["erased"]
    out = os.path.join(HERE, "track_seq%d.npz" % SEED)
    np.savez_compressed(out, **d)
    print(out, os.path.getsize(out), "bytes; created", [r["created"] for r in recs], "erased", [r["erased"] for r in recs])


if __name__ == "__main__":
    main()
