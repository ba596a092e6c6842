#!/usr/bin/env python
"""Per-source-line stall samples from an ncu report:
   ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; ncu_lines.py x.csv [top]"""
import csv, sys
path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur, hdr, out = None, None, []
for r in rows:
    if len(r) == 2 and r[0] == 'File Path':
        cur = r[1].split('/')[-1]
    elif r and r[0] == 'Line No':
        hdr = r
    elif hdr and len(r) == len(hdr) and r[2] == '-':
        d = dict(zip(hdr, r))
        n = int(d['# Samples'] or 0)
        if n:
            stalls = {k[6:]: int(v) for k, v in d.items() if k.startswith('stall_') and '(Not' not in k and v not in ('', '0')}
            best = sorted(stalls.items(), key=lambda x: -x[1])[:3]
            out.append((n, cur, r[0], int(d['Instructions Executed'] or 0), best, r[1].strip()[:90]))
tot = sum(o[0] for o in out)
print('total samples', tot)
for n, f, ln, ex, best, src in sorted(out, key=lambda x: -x[0])[:top]:
    print(f'{100 * n / tot:5.1f}% {f}:{ln} ex={ex} {best} | {src}')
