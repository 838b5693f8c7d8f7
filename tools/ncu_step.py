#!/usr/bin/env python3
"""Summarise one batch step (and one single-frame step) from an ncu launch list (--csv, gpu__time_duration.sum).
A step starts at k_unpack (first kernel of svo_batch_submit) and ends before the next k_unpack."""
import csv
import sys


def steps(rows):
    cur = None
    for r in rows:
        name = r["Kernel Name"].split("(")[0].replace("void ", "")
        if name.startswith("k_unpack"):
            if cur:
                yield cur
            cur = []
        if cur is not None:
            cur.append((name, r["Grid Size"], float(r["Metric Value"].replace(",", "")) / 1000.0))
    if cur:
        yield cur


def nframes(step):
    for n, g, _ in step:
        if n == "k_stereo":
            return int(g.strip("()").split(",")[1])
    return 0


def main(path, want):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    allsteps = list(steps(rows))
    for nf in (want, 1):
        sel = [s for s in allsteps if nframes(s) == nf and any(n.startswith("k_pairs") or n.startswith("k_bf") for n, _, _ in s)]
        if not sel:
            continue
        st = sel[-1]
        if len(sel) > 1 and len(sel[-2]) > len(st):
            st = sel[-2]            # the capture limit cut the last step short
        tot = sum(t for _, _, t in st)
        print("---- step with %d frame(s): %d launches, %.1f us total, %.1f us/frame" % (nf, len(st), tot, tot / nf))
        for n, g, t in st:
            print("%-18s grid=%-16s %9.1f us  %5.1f%%" % (n, g, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32)
