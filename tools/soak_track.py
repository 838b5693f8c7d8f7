#!/usr/bin/env python3
"""Soak test of the device-resident tracker (run on the GPU box): S sequences x T frames at the bench's shape (1241x376,
2000 features, 5000-row capacity) through svo_batch_submit(track_seq) across alternating lanes, every frame's matches
and owned points compared with oracle/track.py, the whole state every 20 frames.  Prints one JSON line.
    python tools/soak_track.py [S] [T]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200"), os.path.join(ROOT, "tests")]
import svo  # noqa: E402
import synth  # noqa: E402
from oracle import track as T  # noqa: E402
import test_gpu_track as G  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
NT = int(sys.argv[2]) if len(sys.argv) > 2 else 120
ctx = svo.Context(1241, 376, nfeatures=2000, max_batch=S, lanes=3, max_rows=5000)
ctx.track_create(S, 5000, 4)
rng = np.random.default_rng(1)
ballast = rng.integers(0, 256, (2500, 32), dtype=np.uint8)
trackers = []
for s in range(S):
    b = ballast if s % 2 else None            # odd sequences carry ballast rows (a fuller map, like the bench)
    ctx.track_reset(s, b)
    trackers.append(T.Tracker(window=4, ballast=b, map_cap=5000))
seqs = [synth.Sequence(synth.K_SHAPE, seed=70 + s) for s in range(S)]
ctx.set_outputs(svo.OUT_COMPACT | svo.OUT_NO_RIGHT)
t0 = time.time()
stats = dict(frames=0, p1=0, p2=0, bad=0, created=0, max_map=0)
for t in range(NT):
    lane = t % 3
    frames = []
    for s in range(S):
        bx = (G.BOXES, G.F_TEST) if (s == 0 and t % 3 == 1) else ((G.BOXES[:1], None) if s == 1 and t % 5 == 2 else (None, None))
        frames.append(G.frame_dict(seqs[s].frame(t), s, t, bx[0], bx[1]))
    ctx.batch_submit(lane, frames); ctx.batch_wait(lane)
    for s in range(S):
        r = ctx.batch_result(lane, s)
        bx = (G.BOXES, G.F_TEST) if (s == 0 and t % 3 == 1) else ((G.BOXES[:1], None) if s == 1 and t % 5 == 2 else (None, None))
        xy = np.stack([r["kp_left"]["x"], r["kp_left"]["y"]], 1)
        o = trackers[s].step(xy, r["desc_left"], r["depth"], t, boxes=bx[0], F=bx[1], K4=G.K4)
        G.compare_frame(r, o, "seq %d frame %d" % (s, t))
        stats["frames"] += 1; stats["created"] += o["created"]; stats["max_map"] = max(stats["max_map"], o["n_map"])
        if o["n_prev"]:
            stats["p1"] += int(o["p1_row_claimed"].sum()); stats["bad"] += int(o["p1_row_bad"].sum())
        if o["n_map"]:
            stats["p2"] += int(o["p2_row_claimed"].sum())
        if t % 20 == 19 or t == NT - 1:
            G.compare_state(ctx.track_state(s), trackers[s], "seq %d after frame %d" % (s, t))
stats["seconds"] = round(time.time() - t0, 1)
stats["sequences"], stats["frames_per_sequence"] = S, NT
print(json.dumps(stats))
ctx.close()
