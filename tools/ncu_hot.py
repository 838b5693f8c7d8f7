#!/usr/bin/env python3
"""Group the SASS of one kernel in a .ncu-rep by execution count to find its hot regions.
usage: ncu_hot.py report.ncu-rep kernel_name [launch_index]"""
import csv
import io
import subprocess
import sys


def main(rep, kernel, which=0):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", kernel],
                         capture_output=True, text=True).stdout
    # several launches are concatenated, each starting with a "Kernel Name" line
    blocks = txt.split('"Kernel Name"')
    blk = '"Kernel Name"' + blocks[1 + which]
    rows = list(csv.reader(io.StringIO(blk)))
    print(rows[0])
    hdr, data = rows[1], [r for r in rows[2:] if len(r) > 10]
    ci, ti, si = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    tot = sum(int(r[ci]) for r in data)
    tots = sum(int(r[si]) for r in data)
    print("total warp instructions", tot, "samples", tots)
    grp, cur = [], None
    for i, r in enumerate(data):
        c = int(r[ci])
        if cur and abs(c - cur[2]) <= 0.02 * max(c, cur[2], 1):
            cur[1] = i; cur[3] += c; cur[4] += int(r[ti]); cur[5] += int(r[si])
        else:
            cur = [i, i, c, c, int(r[ti]), int(r[si])]
            grp.append(cur)
    for g in grp:
        if g[3] > 0.005 * tot or g[5] > 0.005 * tots:
            print("idx %4d-%4d n=%3d exec/instr=%9d warpinst=%5.1f%% thr/inst=%4.1f samples=%5.1f%%  %s"
                  % (g[0], g[1], g[1] - g[0] + 1, g[2], 100 * g[3] / tot, g[4] / max(g[3], 1), 100 * g[5] / max(tots, 1),
                     data[g[0]][1].strip()[:48]))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
