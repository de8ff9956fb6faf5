#!/usr/bin/env python
"""Turn an `ncu --set full` capture (exported with `ncu -i X.ncu-rep --page raw --csv`) into profiles/r01/traffic.json:
per kernel family (the labels bench.py's per-launch profile uses) the measured DRAM bytes per launch
(dram__bytes_read.sum + dram__bytes_write.sum), duration and the tensor/DRAM throughput percentages.
usage: tools/ncu_traffic.py raw.csv [out.json]"""
import csv
import json
import re
import sys


def label(name):
    m = re.search(r'conv3_ws_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+), (?:\(int\))?(\d+)(?:, (?:\(bool\))?(\w+))?>', name)
    if m:
        pair = m.group(4) in ('1', 'true')
        return ('conv1_ws' if m.group(3) != '0' else ('conv3_ws_pair' if pair else 'conv3_ws')) + f'<BN={m.group(1)},CK={m.group(2)}>'
    m = re.search(r'conv_tc_kernel<(?:\(int\))?(\d+), (?:\(int\))?(\d+)>', name)
    if m:
        return f'conv_tc<BN={m.group(1)},BK={m.group(2)}>'
    m = re.search(r'(\w+)_kernel', name)
    return m.group(1) if m else name[:40]


def to_bytes(v, unit):
    v = float(v.replace(',', ''))
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[unit]


def main(path, out):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ix = {k: hdr.index(k) for k in ('Kernel Name', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__time_duration.sum')}
    opt = {k: hdr.index(k) for k in ('sm__pipe_tensor_subunit_cycles_active.avg.pct_of_peak_sustained_elapsed',
                                     'dram__throughput.avg.pct_of_peak_sustained_elapsed',
                                     'sm__throughput.avg.pct_of_peak_sustained_elapsed') if k in hdr}
    agg = {}
    for r in rows[2:]:
        k = label(r[ix['Kernel Name']])
        rd = to_bytes(r[ix['dram__bytes_read.sum']], units[ix['dram__bytes_read.sum']])
        wr = to_bytes(r[ix['dram__bytes_write.sum']], units[ix['dram__bytes_write.sum']])
        du = float(r[ix['gpu__time_duration.sum']].replace(',', ''))
        du *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'usecond': 1e-3, 'msecond': 1.0, 'nsecond': 1e-6}[units[ix['gpu__time_duration.sum']]]
        a = agg.setdefault(k, dict(launches=0, read=0.0, write=0.0, ms=0.0, **{o: 0.0 for o in opt}))
        a['launches'] += 1; a['read'] += rd; a['write'] += wr; a['ms'] += du
        for o, i in opt.items():
            a[o] += float(r[i].replace(',', ''))
    res = {}
    for k, a in agg.items():
        n = a['launches']
        res[k] = dict(launches_captured=n, dram_bytes_per_launch=(a['read'] + a['write']) / n,
                      dram_read_bytes_per_launch=a['read'] / n, dram_write_bytes_per_launch=a['write'] / n,
                      ncu_ms_per_launch=a['ms'] / n, **{o.split('.')[0] + '_pct': a[o] / n for o in opt})
    with open(out, 'w') as f:
        json.dump(res, f, indent=1, sort_keys=True)
    for k, v in sorted(res.items()):
        print(k, json.dumps(v))


if __name__ == '__main__':
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else 'profiles/r01/traffic.json')
