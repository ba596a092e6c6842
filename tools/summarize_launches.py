#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (cold-cache, serialised:
compare SHARES, not absolutes).  usage: summarize_launches.py launches.csv [skip_first_n]"""
import collections
import csv
import sys


def main(path, skip=0):
    hdr, data = None, []
    for r in csv.reader(open(path)):
        if len(r) > 5 and r[0] == 'ID':
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    data = data[skip:]
    agg = collections.OrderedDict()
    for d in data:
        v = float(d['Metric Value'].replace(',', ''))
        v = v / 1000 if d['Metric Unit'] == 'ns' else (v * 1000 if d['Metric Unit'] == 'ms' else v)
        a = agg.setdefault(d['Kernel Name'][:70], [0, 0.0, d['Grid Size'], d['Block Size']])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    print(f'| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|')
    for k, (c, t, g, b) in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f'| `{k}` | {c} | {t:.1f} | {t / c:.1f} | {t / tot:.3f} | {g} | {b} |')
    print(f'\ntotal {tot:.1f} us over {len(data)} launches')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
