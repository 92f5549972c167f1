#!/usr/bin/env python
"""Per-kernel times of the last N launches of an ncu launch list. usage: ncu_tail.py file.csv N"""
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
L = []
for r in rows[1:]:
    m = re.search(r'(\w+Fn|scan_\w+_kernel|sc_\w+_kernel)', r[ki]); nm = m.group(1) if m else r[ki][:40]
    v = float(r[vi].replace(',', '')); u = r[ui]
    v = v / 1000 if u == 'ns' else v * 1000 if u == 'ms' else v
    L.append((nm, v))
n = int(sys.argv[2])
tot = 0
for nm, v in L[-n:]:
    print(f"{nm:22s} {v:8.2f} us"); tot += v
print(f"{'total':22s} {tot:8.2f} us   ({len(L)} launches in the file)")
