#!/usr/bin/env python
"""(dev) Per-barrier-segment attribution of an `ncu --page source --csv` dump: instructions, samples, top stalls,
shared-memory wavefronts.   ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_segments.py src.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
seg = 0; agg = {}
for r in data:
    src = r[ix['Source']].strip()
    op = src.split()[0] if not src.startswith('@') else src.split()[1]
    a = agg.setdefault(seg, dict(inst=0, smp=0, wf=0, st={}))
    a['inst'] += int(r[ix['Instructions Executed']]); a['smp'] += int(r[ix['# Samples']]); a['wf'] += int(r[ix['L1 Wavefronts Shared']])
    for c in h:
        if c.startswith('stall_') and 'Not Issued' not in c:
            v = int(r[ix[c]])
            if v: a['st'][c[6:]] = a['st'].get(c[6:], 0) + v
    if op.startswith('BAR'): seg += 1
ti = sum(a['inst'] for a in agg.values()); tw = sum(a['wf'] for a in agg.values())
print('total warp-inst %d samples %d smem wavefronts %d' % (ti, tot, tw))
for s, a in agg.items():
    if a['smp'] < 20: continue
    top = sorted(a['st'].items(), key=lambda kv: -kv[1])[:5]
    print('seg %2d inst %5.1f%% smp %5.1f%% wf %5.1f%%  %s' % (s, 100 * a['inst'] / ti, 100 * a['smp'] / tot, 100 * a['wf'] / max(tw, 1), top))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
for key in ('stall_long_sb', 'stall_barrier'):
    print('top', key)
    for r in sorted(data, key=lambda r: -int(r[ix[key]]))[:n]:
        print('  %5s %5s %s' % (r[ix[key]], r[ix['# Samples']], r[ix['Source']].strip()[:100]))
