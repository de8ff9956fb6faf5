#!/usr/bin/env python
"""Fused criterion alone at BASELINE config 5 geometry (12 x 1024 x 1024 labels, 19 classes, OHEM 0.9 / 131072): forward and
forward + backward device times (CUDA events).  Used under ncu for the per-kernel instruction counts."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pidnet_b200 import OhemCrossEntropy, BondaryLoss
from pidnet_b200.criterion import FusedCriterion

CW = [0.8373, 0.918, 0.866, 1.0345, 1.0166, 0.9969, 0.9754, 1.0489, 0.8786, 1.0023, 0.9539, 0.9843, 1.1116, 0.9037, 1.0865, 1.0955, 1.0865, 1.1529, 1.0507]

def main():
    dev = torch.device('cuda:0')
    N, H, W, C = int(os.environ.get('BATCH', 12)), 1024, 1024, 19
    g = torch.Generator().manual_seed(11)
    outs = [(2.0 * torch.randn(N, c, H // 8, W // 8, generator=g)).to(dev) for c in (C, C, 1)]
    labels = torch.randint(0, C, (N, H, W), generator=g)
    labels[:, :32, :] = 255
    labels = labels.to(dev)
    bd = (torch.rand(N, H, W, generator=g) > 0.9).float().to(dev)
    crit = FusedCriterion(OhemCrossEntropy(255, 0.9, 131072, torch.tensor(CW)), BondaryLoss())
    reps = int(os.environ.get('REPS', 10))
    res = {}
    for name, grads in (('forward_ms', False), ('forward_backward_ms', True)):
        for _ in range(2):
            crit(outs, labels, bd, need_grads=grads)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            out, _ = crit(outs, labels, bd, need_grads=grads)
        e1.record()
        torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / reps
    res['loss'] = float(out[0])
    print(json.dumps(res))

if __name__ == '__main__':
    main()
