"""Aggregate an ncu source page (sass,cuda view) per source line: samples, executed warp instructions, stall mix.
usage: ncu -i rep.ncu-rep --page source --csv --print-source sass,cuda [--kernel-name ...] > mix.csv; python tools/ncu_lines.py mix.csv [top]"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, kern, hdr, agg = None, None, None, []
stall = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        kern = r[1]
        continue
    if r[0] == 'Kernel Name':
        continue
    if r[0] == 'Line No':
        hdr = r
        ix = {h: i for i, h in enumerate(hdr)}
        continue
    if hdr is None:
        continue
    if r[0] != '':
        def num(v):
            try:
                return int(v)
            except ValueError:
                return 0
        agg.append((kern, cur, int(r[0]), r[1].strip()[:100], num(r[ix['# Samples']]), num(r[ix['Instructions Executed']])))
        for h, i in ix.items():
            if h.startswith('stall_') and 'Not Issued' not in h:
                stall[(kern, h)] += num(r[i])
kernels = sorted({a[0] for a in agg})
for kn in kernels:
    sel = [a for a in agg if a[0] == kn]
    ts, ti = sum(a[4] for a in sel), sum(a[5] for a in sel)
    print('==', kn[:90], 'samples', ts, 'warp-inst', ti)
    byf = collections.Counter()
    for a in sel:
        byf[a[1]] += a[4]
    print('  by file:', {f: round(100 * v / max(ts, 1), 1) for f, v in byf.items()})
    print('  stalls:', {h[6:]: round(100 * v / max(ts, 1), 1) for (k2, h), v in stall.items() if k2 == kn and v > ts * .02})
    for a in sorted(sel, key=lambda a: -a[4])[:top]:
        print(f'  {a[1][:20]:20s} {a[2]:4d} {100 * a[4] / max(ts, 1):5.1f}% inst {a[5]:9d}  {a[3]}')
