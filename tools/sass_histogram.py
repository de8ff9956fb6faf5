#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the shipped library (what proves a Blackwell-native kernel, B200_PROFILING.md):
tcgen05.mma -> UTC*MMA, tcgen05.ld/st -> LDTM/STTM, TMA -> UTMALDG/UTMASTG, legacy tensor path -> HMMA (must be absent).
usage: python tools/sass_histogram.py [lib.so] > profiles/rNN/sass_histogram.md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ['UTCHMMA', 'UTCHMMA.2CTA', 'LDTM', 'STTM', 'UTMALDG', 'UTMASTG', 'UBLKCP', 'SYNCS', 'LDGSTS', 'HMMA', 'ATOMS', 'RED', 'MUFU']


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'pidnet_b200', 'lib', 'libpidnet_b200.so')
    txt = subprocess.run(['cuobjdump', '-sass', lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    funcs = re.split(r'\n\s+Function : ', txt)[1:]
    rows = []
    total = collections.Counter()
    for f in funcs:
        mangled = f.split('\n', 1)[0].strip()
        name = subprocess.run(['c++filt', mangled], stdout=subprocess.PIPE, text=True).stdout.strip()
        name = re.sub(r'pidnet::\(anonymous namespace\)::|pidnet::', '', name)
        name = re.sub(r'\(.*', '', name).replace('void ', '')
        ops = re.findall(r'/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+ )?([A-Z][A-Z0-9_.]*)', f)
        c = collections.Counter()
        for o in ops:
            base = o.split('.')[0]
            c[base] += 1
            if base == 'UTCHMMA' and '.2CTA' in o:
                c['UTCHMMA.2CTA'] += 1
        rows.append((name, len(ops), c))
        total.update(c)
    rows.sort(key=lambda r: (-(r[2]['UTCHMMA']), r[0]))
    print(f'# SASS opcode histogram of `{os.path.relpath(lib, ROOT)}` ({len(rows)} kernels; `cuobjdump -sass`, tools/sass_histogram.py)\n')
    print('Library totals: ' + ', '.join(f'{k} {total[k]}' for k in KEYS) + '\n')
    print('| kernel | instructions | ' + ' | '.join(KEYS) + ' |')
    print('|---|---|' + '---|' * len(KEYS))
    for name, n, c in rows:
        print(f'| `{name}` | {n} | ' + ' | '.join(str(c[k]) if c[k] else '' for k in KEYS) + ' |')


if __name__ == '__main__':
    main()
