#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (no GPU needed):

    python scripts/sass_histogram.py > profiles/r02_sass_histogram.txt
"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pytorch_news_recommender_b200", "libnrms_b200.so")
COLS = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "LDSM", "LDGSTS", "ELECT", "R2UR",
        "RED", "ATOM"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    kernels, cur = [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = collections.Counter()
            kernels.append(cur)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            cur["instrs"] += 1
            cur[m.group(1)] += 1
    print("SASS opcode histogram of pytorch_news_recommender_b200/libnrms_b200.so (sm_100a), per kernel")
    print("made by: scripts/sass_histogram.py (cuobjdump -sass <so>, counted per 'Function :' block); PTX -> SASS names: tcgen05.mma = UTCHMMA,")
    print("tcgen05.commit = UTCBAR, tcgen05.alloc = UTCATOMSWS, tcgen05.ld = LDTM, cp.async.bulk = UBLKCP, mbarrier = SYNCS, mma.sync = HMMA,")
    print("ldmatrix = LDSM, cp.async = LDGSTS, elect.sync = ELECT.  (No UTMALDG: operands are pre-swizzled images fetched with 1-D bulk copies.)")
    print()
    w = max(len(n) for n in names) + 2
    print("kernel".ljust(w) + "  instrs" + "".join(c.rjust(max(len(c), 5) + 2) for c in COLS))
    for name, c in zip(names, kernels):
        print(name.replace("nrms::", "").ljust(w) + str(c["instrs"]).rjust(8) + "".join(str(c[k]).rjust(max(len(k), 5) + 2) for k in COLS))


if __name__ == "__main__":
    main()
