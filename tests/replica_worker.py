"""Worker for tests/test_replicas_gloo.py: one process per (pretend) GPU, gloo backend, launched by torchrun."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "stereo-semantic-vo_b200"))

import replicas  # noqa: E402
import synth  # noqa: E402


def main():
    out_dir = sys.argv[1]
    n_seq = int(sys.argv[2])
    g = replicas.Group(backend="gloo")
    mine = replicas.assign_sequences(n_seq, g.world, g.rank)
    # every rank synthesises only its own sequences (seed = sequence id), as bench.py does
    checks = {}
    for s in mine:
        L, R = synth.Sequence((96, 160), seed=s).frame(1)
        checks[s] = int(L.astype(np.int64).sum() * 31 + R.astype(np.int64).sum())
    g.barrier()
    ms = 10.0 * (g.rank + 1)                       # pretend device time: rank 1 is the slowest of two
    frames = 100 * len(mine)
    fps, tot_frames, tot_ms = g.aggregate_fps(frames, ms)
    mx = g.max_over_ranks([ms, 1.0])
    g.barrier()
    json.dump({"rank": g.rank, "world": g.world, "mine": mine, "checks": checks, "fps": fps, "frames": tot_frames,
               "ms": tot_ms, "mx": mx, "backend": g.backend}, open(os.path.join(out_dir, "rank%d.json" % g.rank), "w"))
    g.close()


if __name__ == "__main__":
    main()
