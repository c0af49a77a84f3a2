#!/usr/bin/env python
"""Top stall sites of a kernel from an `ncu --page source --csv` export: python tools/ncu_hot.py FILE [N]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
h = rows[1]
ix = {c: i for i, c in enumerate(h)}
body = [r for r in rows[2:] if len(r) == len(h)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot)
stalls = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
agg = {c: sum(int(r[ix[c]] or 0) for r in body) for c in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for k, r in sorted(enumerate(body), key=lambda kr: -int(kr[1][ix["# Samples"]]))[:n]:
    top = sorted(((int(r[ix[c]] or 0), c) for c in stalls), reverse=True)[:2]
    print(f"{k:5d} {int(r[ix['# Samples']]):6d} {r[ix['Source']].strip()[:70]:70s} {top}")
