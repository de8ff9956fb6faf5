#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (markdown)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, per_step_launches=None):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        name = r['Kernel Name']
        m = re.search(r'(\w+_kernel)(<[^>]*>)?', name)
        short = (m.group(1) + (m.group(2) or '')) if m else name[:60]
        short = short.replace('(int)', '')
        rows.append((short, r['Grid Size'], float(r['Metric Value']) / 1e3))
    if per_step_launches:
        rows = rows[:per_step_launches]
    tot = sum(t for _, _, t in rows)
    agg = OrderedDict()
    for k, _, t in rows:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += t
    print(f'launches: {len(rows)}  total device time: {tot / 1e3:.3f} ms (ncu per-launch times are cold-cache and serialised)\n')
    print('| kernel | launches | total us | share |')
    print('|---|---:|---:|---:|')
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f'| `{k}` | {n} | {t:.1f} | {100 * t / tot:.1f} % |')
    print('\n| # | kernel | grid | us |')
    print('|---:|---|---|---:|')
    for i, (k, g, t) in enumerate(rows):
        print(f'| {i} | `{k}` | {g} | {t:.1f} |')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)
