#!/usr/bin/env python
"""Markdown table of selected metrics from `ncu -i X.ncu-rep --page raw --csv` exports (one row per captured launch).
usage: tools/ncu_full_summary.py raw1.csv [raw2.csv ...] > profiles/rNN/ncu_full_*.md"""
import csv
import re
import sys

COLS = [('gpu__time_duration.sum', 'duration'),
        ('sm__cycles_elapsed.avg', 'elapsed cycles / SM'),
        ('l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'L1/smem pipe % of peak'),
        ('dram__bytes_read.sum', 'dram read'), ('dram__bytes_write.sum', 'dram write'),
        ('dram__throughput.avg.pct_of_peak_sustained_elapsed', 'dram % of peak'),
        ('lts__t_sector_hit_rate.pct', 'L2 hit %'), ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm throughput %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue active %'),
        ('launch__registers_per_thread', 'regs'), ('launch__grid_size', 'grid'), ('launch__cluster_size', 'cluster')]


def short(name):
    m = re.search(r'(\w+_kernel)(<[^>]*>)?', name)
    return ((m.group(1) + (m.group(2) or '')) if m else name[:60]).replace('(int)', '').replace('(bool)', '')


def main(paths):
    print('| # | kernel | ' + ' | '.join(c[1] for c in COLS) + ' |')
    print('|---|---|' + '---|' * len(COLS))
    n = 0
    for path in paths:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            cells = []
            for key, _ in COLS:
                if key not in hdr:
                    cells.append('-')
                    continue
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f'{float(v.replace(",", "")):.4g}'
                except ValueError:
                    pass
                u = units[i]
                cells.append(f'{v} {u}'.strip() if u not in ('', '%') else v + (' %' if u == '%' else ''))
            print(f'| {n} | `{short(r[hdr.index("Kernel Name")])}` | ' + ' | '.join(cells) + ' |')
            n += 1


if __name__ == '__main__':
    main(sys.argv[1:])
