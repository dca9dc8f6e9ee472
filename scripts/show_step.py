#!/usr/bin/env python
"""Print the kernels of the last full train step in an ncu launch-list CSV (scripts/launch_list.sh)."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
seq = [(r[kn].split("(")[0].replace("void ", "").replace("nrms::", "")[:60], float(r[mv].replace(",", "")) / 1000)
       for r in rows[hi + 1:] if len(r) > mv]
idx = [i for i, (n, _) in enumerate(seq) if "adam" in n]
a, b = idx[-3], idx[-1]
tot = 0.0
for n, t in seq[a + 1:b + 1]:
    tot += t
    print(f"{t:8.1f} us  {n}")
print(f"total {tot:.1f} us")
