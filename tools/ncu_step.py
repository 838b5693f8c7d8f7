#!/usr/bin/env python3
"""Summarise one batch step (and one single-frame step) from an ncu launch list (--csv, gpu__time_duration.sum)."""
import csv
import sys


def main(path, nframes):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    for want in (nframes, 1):
        last = None
        for i, r in enumerate(rows):
            g = r["Grid Size"].strip("()").split(",")
            if names[i].startswith("k_bf(") and int(g[1]) == want:
                last = i
        if last is None:
            continue
        start = max(i for i in range(1, last) if names[i].startswith("k_resize") and not names[i - 1].startswith("k_resize"))
        tot = 0.0
        out = []
        for r in rows[start:start + 25]:
            t = float(r["Metric Value"].replace(",", "")) / 1000.0
            tot += t
            out.append((r["Kernel Name"].split("(")[0], r["Grid Size"], t))
        print("---- step with %d frame(s): %d launches, %.1f us total, %.1f us/frame" % (want, len(out), tot, tot / want))
        for n, g, t in out:
            print("%-18s grid=%-16s %9.1f us  %5.1f%%" % (n, g, t, 100 * t / tot))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 32)
