#!/usr/bin/env python
"""Golden vectors for the criterion from the LIVE reference (build container only): the unmodified
`utils.utils.FullModel` + `utils.criterion.{OhemCrossEntropy,BondaryLoss}` (imported through a yacs shim,
config set to the YAML values) on seeded logits/labels; stores loss, loss_s, loss_b, acc and the autograd
gradients of losses.mean() w.r.t. the three low-res logit maps."""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


class CN(dict):
    def __getattr__(s, k):
        try:
            return s[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(s, k, v):
        s[k] = v

    def defrost(s): pass
    def freeze(s): pass
    def merge_from_file(s, f): pass
    def merge_from_list(s, l): pass


m, mc = types.ModuleType('yacs'), types.ModuleType('yacs.config')
mc.CfgNode = CN
m.config = mc
sys.modules['yacs'], sys.modules['yacs.config'] = m, mc
sys.path.insert(0, '/root/reference')

import numpy as np  # noqa: E402
import torch  # noqa: E402
from configs import config  # noqa: E402
from utils.criterion import BondaryLoss, OhemCrossEntropy  # noqa: E402
from utils.utils import FullModel  # noqa: E402

from oracle import criterion_oracle as CO  # noqa: E402

CASES = [  # n, ncls, h, w, seed, min_kept, logit scale, class weights, aligned
    (2, 19, 64, 128, 1, 131072, 3.0, True, False),    # k = n-1: threshold = max p of valid pixels vs 0.9
    (2, 19, 64, 128, 2, 500, 3.0, True, False),       # many hard pixels: threshold stays 0.9
    (2, 19, 128, 128, 4, 4000, 2.0, True, True),   # confident logits: k-th order statistic > 0.9 (radix select path)
    (1, 11, 96, 160, 5, 3000, 6.0, False, False),     # CamVid-like: 11 classes, no class weights
]


class Dummy(torch.nn.Module):
    def __init__(self, outs):
        super().__init__()
        self.o = outs

    def forward(self, x):
        return list(self.o)


def main():
    config.LOSS.BALANCE_WEIGHTS = [0.4, 1.0]
    config.LOSS.SB_WEIGHTS = 1.0
    config.LOSS.OHEMTHRES = 0.9
    config.TRAIN.IGNORE_LABEL = 255
    out_dir = os.path.join(ROOT, 'tests', 'golden')
    for (n, ncls, h, w, seed, keep, scale, wts, aligned) in CASES:
        outs, labels, bd = CO.synthetic_batch(n, ncls, h, w, seed, scale=scale, aligned=aligned)
        outs = [o.requires_grad_(True) for o in outs]
        weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS) if wts else None
        config.LOSS.OHEMKEEP = keep
        fm = FullModel(Dummy(outs), OhemCrossEntropy(ignore_label=255, thres=0.9, min_kept=keep, weight=weight), BondaryLoss())
        losses, ups, acc, ll = fm(torch.zeros(1), labels, bd)
        with torch.no_grad():
            pm = torch.softmax(ups[1], 1).gather(1, labels.clamp(max=ncls - 1).unsqueeze(1))[:, 0][labels != 255]
            print('   valid', pm.numel(), 'p<0.9:', int((pm < 0.9).sum()), 'k', min(keep, pm.numel() - 1))
        loss = losses.mean()
        loss.backward()
        rec = dict(n=n, ncls=ncls, h=h, w=w, seed=seed, min_kept=keep, scale=scale, weights=wts, aligned=aligned,
                   loss=loss.item(), loss_s=ll[0].mean().item(), loss_b=ll[1].item(), acc=acc.item(),
                   g_p=outs[0].grad.numpy(), g_m=outs[1].grad.numpy(), g_d=outs[2].grad.numpy(),
                   x_p=outs[0].detach().numpy(), x_m=outs[1].detach().numpy(), x_d=outs[2].detach().numpy(),
                   labels=labels.numpy().astype(np.uint8), bd=bd.numpy().astype(np.uint8))
        fn = os.path.join(out_dir, f'criterion_c{ncls}_{h}x{w}_k{keep}_s{seed}.npz')
        np.savez_compressed(fn, **rec)
        print(fn, os.path.getsize(fn) // 1024, 'KiB', 'loss', loss.item())


if __name__ == '__main__':
    main()
