"""Records what the REFERENCE'S OWN CODE computes for two stereo frames driven like Tracking::Track /
Tracklastframe (src/Tracking.cc:184-250, :114) with offline YOLO boxes, as golden fixtures for the GPU tests of the
C++ drop-in adapter (tests/test_gpu_adapter.py) — /root/reference does not exist on the GPU box.

The reference runs as oracle/_ref/libsvo_ref.so: src/pnpmatch.cc, src/frame.cc and src/mappoint.cc compiled unmodified
(oracle/Makefile `ref`), OpenCV calls answered by cv2 (oracle/ref.py).  Inputs are NOT stored: they are regenerated
from the seeds by synth.Sequence / synth.dense_disparity (a checksum of each is stored instead).

    python tests/golden/make_golden_ref.py        # writes tests/golden/ref_track_seed{5,9}.npz
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200")]
import synth  # noqa: E402
from oracle import ref as R  # noqa: E402

CAL = synth.KITTI_04_12
K = np.array([[CAL["fx"], 0, CAL["cx"]], [0, CAL["fy"], CAL["cy"]], [0, 0, 1]], np.float32)
BF = np.float32(CAL["bf"])
CASES = {5: [[300, 700, 100, 300], [900, 1100, 50, 200], [20, 180, 200, 360]], 9: [[500, 1241, 0, 376]]}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def inputs(seed, shape=synth.K_SHAPE):
    seq = synth.Sequence(shape, seed=seed)
    frames = (seq.frame(0), seq.frame(1))
    disps = (synth.dense_disparity(shape, 2 * seed), synth.dense_disparity(shape, 2 * seed + 1))
    return frames, disps


def main():
    for seed, boxes in CASES.items():
        frames, disps = inputs(seed)
        run = R.run_two_frames(frames, disps, K, BF, boxes)
        cur, last, f0, map0 = run["cur"], run["last"], run["f0"], run["before"]["map"]
        assert f0["N"] == len(f0["kps"]) == 500 and cur["N"] == len(cur["kps"]) == 500, "pick a seed with exactly 500 keypoints"
        out = os.path.join(HERE, "ref_track_seed%d.npz" % seed)
        np.savez_compressed(
            out, seed=seed, boxes=np.asarray(boxes, np.int32), K=K, bf=BF,
            input_sha=np.array([sha(frames[0][0]), sha(frames[0][1]), sha(frames[1][0]), sha(frames[1][1]), sha(disps[0]), sha(disps[1])]),
            f0_kps=f0["kps"], f0_desc=f0["desc"], f0_keypoints_r=f0["keypoints_r"], f0_depth_at_kp=f0["depth_at_kp"],
            map_idx=map0["idx"], map_worldpos=map0["worldpos"], map_desc=map0["desc"], created=run["created"],
            F=run["F"]["F"], F_p1=run["F"]["p1"], F_p2=run["F"]["p2"],
            cur_kps=cur["kps"], cur_desc=cur["desc"], cur_match_score=cur["match_score"],
            cur_mp_idx=np.where(cur["mp_create_id"] == 0, cur["mp_idx"], -1), last_mp_bad=last["mp_bad"],
            last_has_mp=(run["before"]["last"]["mp_create_id"] >= 0).astype(np.uint8),
            pnp_p3=run["pnp"]["p3"], pnp_p2=run["pnp"]["p2"], cur_Tcw=cur["Tcw"])
        print(out, "map points", run["created"], "matched", int((cur["mp_create_id"] >= 0).sum()), "bad", int(last["mp_bad"].sum()))


if __name__ == "__main__":
    main()
