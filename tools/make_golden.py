#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ from the LIVE reference (/root/reference).

Run in the build container only (the reference does not exist on the GPU box):
    python tools/make_golden.py
For every case the UNMODIFIED reference `models.pidnet.PIDNet` is instantiated with the reference's
own constructor, loaded with a seeded state_dict (oracle.make_state_dict -- regenerable anywhere
from the seed; its sha256 is stored so a different RNG stream is detected, not silently accepted),
run on a seeded input in eval mode on CPU fp32, and the outputs are stored.  One tiny case also
stores the full weights so it is independent of the RNG.
"""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
sys.path.insert(0, '/root/reference')

import numpy as np  # noqa: E402
import torch  # noqa: E402

import models.pidnet as REF  # noqa: E402  (the reference)
from oracle import pidnet_oracle as O  # noqa: E402

CASES = [
    # name, classes, augment, N, H, W, weight seed, input seed, store weights
    ('tiny_s', 5, True, 2, 64, 128, 101, 201, True),
    ('tiny_l', 5, True, 1, 128, 128, 102, 202, False),
    ('s', 19, False, 1, 128, 256, 103, 203, False),
    ('s', 19, True, 1, 64, 64, 104, 204, False),
    ('m', 11, True, 1, 120, 160, 105, 205, False),     # CamVid aspect, odd pooled sizes
    ('l', 19, True, 1, 64, 128, 106, 206, False),
]


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def main():
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(out_dir, exist_ok=True)
    torch.set_num_threads(4)
    for name, ncls, aug, N, H, W, wseed, xseed, store in CASES:
        cfg = O.config_for(name, ncls, aug)
        sd = O.make_state_dict(cfg, wseed)
        ref = REF.PIDNet(m=cfg['m'], n=cfg['n'], num_classes=ncls, planes=cfg['planes'],
                         ppm_planes=cfg['ppm_planes'], head_planes=cfg['head_planes'], augment=aug).eval()
        ref.load_state_dict(sd, strict=True)
        x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(xseed))
        with torch.no_grad():
            y = ref(x)
        ys = y if aug else [y]
        rec = dict(name=name, num_classes=ncls, augment=aug, wseed=wseed, xseed=xseed, sd_sha256=sd_digest(sd),
                   x=x.numpy(), n_out=len(ys))
        for i, t in enumerate(ys):
            rec[f'out{i}'] = t.numpy()
        if store:
            for k, v in sd.items():
                rec['w::' + k] = v.numpy()
        fn = os.path.join(out_dir, f'pidnet_{name}_c{ncls}_a{int(aug)}_{H}x{W}.npz')
        np.savez_compressed(fn, **rec)
        print(fn, os.path.getsize(fn) // 1024, 'KiB')


if __name__ == '__main__':
    main()
