#!/usr/bin/env python
"""Top stall sites of an `ncu --page source --csv` dump: python tools/ncu_hot.py file.csv [n]"""
import csv
import sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
samp = col['# Samples']
stalls = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
body = [r for r in rows[2:] if len(r) > samp and r[samp].isdigit()]
tot = sum(int(r[samp]) for r in body)
print('total samples', tot)
agg = {s: sum(int(r[col[s]]) for r in body if r[col[s]].isdigit()) for s in stalls}
print('by reason:', ', '.join('%s %.1f%%' % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(body, key=lambda r: -int(r[samp]))[:n]:
    top = sorted(((int(r[col[s]]), s[6:]) for s in stalls if r[col[s]].isdigit()), reverse=True)[:2]
    print('%6d %5.1f%%  %-70s %s' % (int(r[samp]), 100.0 * int(r[samp]) / tot, r[col['Source']].strip()[:70], top))
