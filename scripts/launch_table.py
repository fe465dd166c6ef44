#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel.  Usage: launch_table.py file.csv [--seq]"""
import collections
import csv
import re
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, data = None, []
    for r in rows:
        if 'Kernel Name' in r:
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            data.append(dict(zip(hdr, r)))
    out = []
    for d in data:
        name = re.sub(r'\(.*', '', d['Kernel Name']).replace('void ', '').replace('tok::', '')
        v = float(d['Metric Value'].replace(',', ''))
        unit = d['Metric Unit']
        us = v / 1e3 if unit in ('ns', 'nsecond') else (v * 1e3 if unit in ('ms', 'msecond') else v)
        out.append((name, d['Grid Size'], us))
    return out


def main():
    data = load(sys.argv[1])
    if '--seq' in sys.argv:
        for i, (n, g, us) in enumerate(data):
            print(i, n[:44].ljust(44), g.ljust(16), f'{us:.1f}')
        return
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, g, us in data:
        agg[n][0] += 1
        agg[n][1] += us
    tot = sum(v[1] for v in agg.values())
    print('| kernel | launches | total us | share |\n|---|---|---|---|')
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| {k[:80]} | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |')
    print(f'| **total** | {len(data)} | {tot:.1f} | 100% |')


if __name__ == '__main__':
    main()
