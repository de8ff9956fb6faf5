"""Print a per-launch profile written by bench.py --profile-out, optionally next to an older one."""
import json, sys
new = json.load(open(sys.argv[1]))
old = {o['name']: o for o in json.load(open(sys.argv[2]))} if len(sys.argv) > 2 else {}
flt = sys.argv[3] if len(sys.argv) > 3 else ''
tot_n = sum(o['ms'] for o in new)
print(f'serialised sum {tot_n:.3f} ms' + (f" (old {sum(o['ms'] for o in old.values()):.3f})" if old else ''))
for o in new:
    if flt and flt not in o['name'] and flt not in o['kernel']:
        continue
    om = old.get(o['name'], {}).get('ms')
    gbs = o['bytes'] / o['ms'] / 1e6
    tf = o['flops'] / o['ms'] / 1e9
    d = f'  was {om:.4f}' if om is not None and abs(om - o["ms"]) > 0.003 else ''
    print(f"{o['name']:28s} {o['kernel']:24s} {o['ms']:.4f} ms {gbs:7.0f} GB/s {tf:7.0f} TF/s{d}")
