#!/usr/bin/env python3
"""Prints the in-kernel clock64 timeline of CTA (0, 0) of the tensor-core matcher kernels (csrc/tcham.cu) for one
bench-like batch.  Run on the GPU box:  SVO_B200_TC_PROF=1 python tools/tc_timeline.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "stereo-semantic-vo_b200")]
os.environ.setdefault("SVO_B200_TC_PROF", "1")
import svo  # noqa: E402
import synth  # noqa: E402

B = 32
ctx = svo.Context(1241, 376, nfeatures=2000, max_batch=B, lanes=1, max_rows=5000)
seq = synth.Sequence(seed=0)
frames = [seq.frame(t) for t in range(3)]
descs = [ctx.extract(f[0])[1] for f in frames]
mp = synth.local_map(descs[:2], rows=5000, seed=1)
jobs = [dict(left=frames[2][0], right=frames[2][1], bf=379.8, baseline=0.537, prev_desc=descs[1], map_desc=mp) for _ in range(B)]
for _ in range(3):
    ctx.batch_submit(0, jobs); ctx.batch_wait(0)
st = ctx.tc_profile()
for mode, name in enumerate(("pairs", "scores", "short")):
    s = st[mode]
    t0 = s[3, 0]
    if t0 == 0:
        continue
    print("== %s: CTA(0,0) timeline, cycles since kernel entry" % name)
    print("  setup done %d | roles done %d | after barrier %d | exit %d" % tuple(int(s[3, k] - t0) for k in (1, 2, 3, 4)))
    print("  producer issue times :", [int(x - t0) for x in s[0] if x][:24])
    m = [int(x - t0) for x in s[1] if x]
    print("  mma (tempty ok, full ok) pairs:", list(zip(m[0::2], m[1::2]))[:24])
    e = [int(x - t0) if x else None for x in s[2]][:30]
    for i in range(0, len(e), 6):
        if e[i] is None:
            break
        print("  epilogue warp 0 tile %d: wait from %s, tfull at %s, chunks at %s" % (i // 6, e[i], e[i + 1], e[i + 2:i + 6]))
ctx.close()
