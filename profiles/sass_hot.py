#!/usr/bin/env python
"""Top SASS instructions by stall samples from `ncu --page source --csv --print-source sass`."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Source' in r and len(r) > 5)
hdr = rows[hi]; ix = {h: i for i, h in enumerate(hdr)}
data = []
for r in rows[hi + 1:]:
    if len(r) != len(hdr) or r == hdr:
        if data:
            break          # first kernel instance only
        continue
    data.append(r)
tot = sum(int(r[ix['# Samples']] or 0) for r in data)
tot_inst = sum(int(r[ix['Instructions Executed']] or 0) for r in data)
print('total samples', tot, 'warp instructions', tot_inst)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']] or 0))[:n]:
    st = sorted(((int(r[ix[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print(f"{int(r[ix['# Samples']]):5d} {100*int(r[ix['# Samples']])/max(tot,1):5.1f}%  exec={r[ix['Instructions Executed']]:>8}  {st}  {r[ix['Source']][:90]}")
