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
    global COLS
    # tensor-pipe activity (tcgen05 MMAs): whichever of ncu's tensor-pipe metrics this capture holds
    hdr0 = next(csv.reader(open(paths[0])))
    want = [('sm__ops_path_tensor_op_hmma_src_bf16_dst_fp32_sparsity_off.avg.pct_of_peak_sustained_elapsed', 'tensor ops bf16->fp32 % of peak'),
            ('TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed', 'tensor pipe active % (elapsed)'),
            ('sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active', 'tensor hmma subpipe inst % (active)')]
    COLS = COLS[:3] + [w for w in want if w[0] in hdr0] + COLS[3:]
    derived = 'TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg' in hdr0
    print('| # | kernel | ' + ' | '.join(c[1] for c in COLS) + (' | tensor sub-pipe cycles active / (4 x elapsed) |' if derived else ' |'))
    print('|---|---|' + '---|' * (len(COLS) + (1 if derived else 0)))
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
            # tcgen05 activity: ncu reports it on the hmma sub-pipe counter (cycles summed over the SM's four tensor sub-pipes)
            key_t, key_e = 'TPC.TriageCompute.sm__pipe_tensor_subpipe_hmma_cycles_active_realtime.avg', 'sm__cycles_elapsed.avg'
            if key_t in hdr and key_e in hdr:
                try:
                    t = float(r[hdr.index(key_t)].replace(',', '')); e = float(r[hdr.index(key_e)].replace(',', ''))
                    cells.append(f'{100 * t / (4 * e):.1f} %')
                except ValueError:
                    cells.append('-')
            print(f'| {n} | `{short(r[hdr.index("Kernel Name")])}` | ' + ' | '.join(cells) + ' |')
            n += 1


if __name__ == '__main__':
    main(sys.argv[1:])
