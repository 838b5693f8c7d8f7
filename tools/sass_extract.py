#!/usr/bin/env python3
"""profiles/r2_sass_extract.txt: what the shipped libsvo_b200.so contains, per kernel, from `cuobjdump -sass`
(no GPU needed): instruction count and the counts of the mnemonics that prove the Blackwell paths —
UTCIMMA / UTCBAR / LDTM (tcgen05.mma.kind::i8, tcgen05.commit, tcgen05.ld), UBLKCP (cp.async.bulk), SYNCS (mbarrier),
VIMNMX3 / VIMNMX (DPX packed min/max), POPC, FMUL2 / FADD2 (packed fp32), REDUX — plus the first lines that hold each.
usage: sass_extract.py [lib] > profiles/r2_sass_extract.txt"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "stereo-semantic-vo_b200/libsvo_b200.so"
KEYS = ["UTCIMMA", "UTCBAR", "LDTM", "UTCATOMSWS", "UBLKCP", "UTMALDG", "SYNCS", "VIMNMX3", "VIMNMX", "POPC", "FMUL2", "FADD2", "FFMA2",
        "REDUX", "MATCH", "SHFL", "BAR.SYNC", "LDS", "STS", "LDG", "STG", "ATOMS", "ATOMG", "RED"]
txt = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
fn = None
count = collections.OrderedDict()
first = {}
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().replace("(anonymous namespace)::", "").replace("void ", "").split("(")[0]
        count[fn] = collections.Counter()
        first[fn] = {}
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4})\*/\s+(.*?);", line)
    if fn and m:
        ins = m.group(2)
        ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
        op = ins.split()[0]
        count[fn]["_total"] += 1
        for k in KEYS:
            if op == k or op.startswith(k + "."):
                if k == "VIMNMX" and op.startswith("VIMNMX3"):
                    continue
                count[fn][k] += 1
                first[fn].setdefault(k, "/*%s*/ %s" % (m.group(1), ins))
print("libsvo_b200.so (sm_100a) — per kernel: SASS instructions, then mnemonic counts of interest\n")
for fn, c in count.items():
    if not fn.startswith("k_") and "k_" not in fn:
        continue
    keys = [k for k in KEYS if c[k]]
    print("%-44s %6d instr  %s" % (fn, c["_total"], "  ".join("%s=%d" % (k, c[k]) for k in keys)))
print("\nfirst occurrence of the Blackwell-specific mnemonics per kernel\n")
for fn, c in count.items():
    for k in ("UTCIMMA", "UTCBAR", "LDTM", "UBLKCP", "SYNCS", "VIMNMX3", "FMUL2", "FADD2"):
        if k in first[fn]:
            print("%-34s %s" % (fn[:34], first[fn][k]))
