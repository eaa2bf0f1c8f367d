#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu --page source --csv`."""
import csv, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
# a report with several kernels repeats the header row per kernel: keep the instruction rows only (first kernel's columns)
data = [r for r in rows[2:] if len(r) == len(hdr) and r != hdr and (r[ix["# Samples"]] or "0").isdigit()]
tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
order = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]] or 0))[:topn]
for i in sorted(order):
    r = data[i]
    s = int(r[ix["# Samples"]] or 0)
    st = sorted(((int(r[ix[c]] or 0), c[6:]) for c in stall_cols), reverse=True)[:3]
    print(f"{i:5d} {100*s/tot:5.1f}% ex={r[ix['Instructions Executed']]:>8s} {r[ix['Source']].strip()[:70]:70s} {' '.join(f'{n}:{v}' for v,n in st if v)}")
