#!/usr/bin/env python
"""Region view of an `ncu --page source --csv` export: contiguous SASS runs with the same
execution count (= same loop nest), with their share of samples and of executed instructions,
plus the kernel's stall-reason totals.   python scripts/ncu_regions.py src.csv [items]"""
import collections
import csv
import sys

csv.field_size_limit(10**9)
rows = list(csv.reader(open(sys.argv[1])))
items = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
kern, hdr, data = None, None, {}
for r in rows:
    if r and r[0] == 'Kernel Name':
        kern = r[1]
        if kern in data:
            kern = None
        else:
            data[kern] = []
        continue
    if r and r[0] == 'Address':
        hdr = r
        continue
    if kern and hdr and len(r) >= 10 and r[0].startswith('0x'):
        data[kern].append(r)
for k, body in data.items():
    cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    tot_s = sum(int(r[2]) for r in body)
    print(k[:70], '| SASS', len(body), '| samples', tot_s)
    print('   stalls:', {hdr[i][6:]: sum(int(r[i] or 0) for r in body) for i in cols if sum(int(r[i] or 0) for r in body) > 0.03 * tot_s})
    runs = []
    for r in body:
        c, smp = int(r[5]), int(r[2])
        t = r[1].strip()
        op = (t.split()[1] if t.startswith('@') else t.split()[0]).split('.')[0]
        if runs and runs[-1][0] == c:
            runs[-1][1] += 1
            runs[-1][2][op] += 1
            runs[-1][3] += smp
        else:
            runs.append([c, 1, collections.Counter({op: 1}), smp])
    tot = sum(c * n for c, n, _, _ in runs)
    print(f'   warp-instr per item {tot / items:.0f}')
    for c, n, ops, smp in sorted(runs, key=lambda x: -x[3])[:10]:
        print(f"   samples {100 * smp / tot_s:5.1f}%  instr {100 * c * n / tot:5.1f}%  count/item {c / items:7.2f} x {n:4d} instrs  {dict(ops.most_common(6))}")
