#!/usr/bin/env python
"""Summarise an `ncu --page source --csv` dump: top SASS instructions by stall samples, grouped by opcode/reason."""
import csv, sys, collections
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
data = rows[2:]
tot = sum(float(r[ix['# Samples']] or 0) for r in data)
print('total samples', tot, 'instructions', len(data))
agg = collections.Counter()
for r in data:
    for h in stall_cols:
        agg[h] += float(r[ix[h]] or 0)
print('by reason:', [(h, int(v)) for h, v in agg.most_common(8)])
byop = collections.Counter(); cnt = collections.Counter(); ex = collections.Counter()
for r in data:
    op = r[ix['Source']].split()[0] if r[ix['Source']].split() else '?'
    if op.startswith('@'): op = r[ix['Source']].split()[1]
    byop[op] += float(r[ix['# Samples']] or 0); cnt[op] += 1; ex[op] += float(r[ix['Instructions Executed']] or 0)
print('by opcode (samples, static count, executed):')
for op, v in byop.most_common(14): print(f'  {op:22s} {int(v):8d} {cnt[op]:6d} {int(ex[op]):12d}')
print('top instructions:')
for r in sorted(data, key=lambda r: -float(r[ix['# Samples']] or 0))[:top]:
    reasons = sorted(((h, float(r[ix[h]] or 0)) for h in stall_cols), key=lambda t: -t[1])[:3]
    print(f"  {r[ix['Address']][-5:]} {int(float(r[ix['# Samples']])):6d} {r[ix['Source']][:70]:70s} {[(h[6:], int(v)) for h, v in reasons if v]}")
