"""CPU tests pinning the criterion oracle: golden vectors generated from the live reference criterion
(tools/make_golden_criterion.py) and, where /root/reference is mounted, the live reference itself."""
import glob
import os
import sys
import types

import numpy as np
import pytest
import torch

from oracle import criterion_oracle as CO

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'criterion_*.npz')))


def load(path):
    z = np.load(path)
    outs = [torch.from_numpy(z[k]).clone() for k in ('x_p', 'x_m', 'x_d')]
    labels = torch.from_numpy(z['labels'].astype(np.int64))
    bd = torch.from_numpy(z['bd'].astype(np.float32))
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS) if bool(z['weights']) else None
    return z, outs, labels, bd, weight, dict(ohem_keep=int(z['min_kept']))


def test_golden_exist():
    assert len(GOLDEN) >= 4


@pytest.mark.parametrize('path', GOLDEN, ids=os.path.basename)
def test_oracle_matches_reference_golden(path):
    z, outs, labels, bd, weight, cfg = load(path)
    outs = [o.requires_grad_(True) for o in outs]
    losses, ups, acc, ll = CO.full_model_forward(outs, labels, bd, weight, cfg)
    loss = losses.mean()
    loss.backward()
    assert losses.shape == (1,) + tuple(labels.shape)            # the reference's [1,N,H,W] quirk
    assert abs(loss.item() - float(z['loss'])) <= 1e-6 * abs(float(z['loss']))
    assert abs(ll[0].mean().item() - float(z['loss_s'])) <= 1e-6 * abs(float(z['loss_s']))
    assert abs(ll[1].item() - float(z['loss_b'])) <= 1e-6 * abs(float(z['loss_b']))
    assert abs(acc.item() - float(z['acc'])) < 1e-7
    for g, k in zip(outs, ('g_p', 'g_m', 'g_d')):
        assert np.abs(g.grad.numpy() - z[k]).max() <= 1e-7 + 1e-5 * np.abs(z[k]).max()


def test_empty_ohem_set_raises_like_reference():
    outs, labels, bd = CO.synthetic_batch(1, 19, 64, 64, 3, scale=0.1)      # no boundary logit passes sigmoid > 0.8
    with pytest.raises(IndexError):
        CO.full_model_forward(outs, labels, bd, None, {})


@pytest.mark.skipif(not os.path.isdir('/root/reference/utils'), reason='reference not mounted')
def test_oracle_matches_live_reference_criterion():
    class CN(dict):
        __getattr__ = lambda s, k: s[k]
        __setattr__ = lambda s, k, v: s.__setitem__(k, v)
        def defrost(s): pass
        def freeze(s): pass
        def merge_from_file(s, f): pass
        def merge_from_list(s, l): pass
    if 'yacs' not in sys.modules:
        m, mc = types.ModuleType('yacs'), types.ModuleType('yacs.config')
        mc.CfgNode = CN
        m.config = mc
        sys.modules['yacs'], sys.modules['yacs.config'] = m, mc
    sys.dont_write_bytecode = True
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
    from configs import config
    from utils.criterion import BondaryLoss, OhemCrossEntropy
    from utils.utils import FullModel
    config.LOSS.BALANCE_WEIGHTS = [0.4, 1.0]
    config.LOSS.SB_WEIGHTS = 1.0
    config.LOSS.OHEMTHRES = 0.9
    config.TRAIN.IGNORE_LABEL = 255

    class Dummy(torch.nn.Module):
        def __init__(self, o):
            super().__init__()
            self.o = o
        def forward(self, x):
            return list(self.o)
    for keep, seed, aligned in [(131072, 11, False), (700, 12, False), (2000, 13, True)]:
        outs, labels, bd = CO.synthetic_batch(2, 19, 64, 96, seed, aligned=aligned)
        weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
        config.LOSS.OHEMKEEP = keep
        fm = FullModel(Dummy(outs), OhemCrossEntropy(255, 0.9, keep, weight), BondaryLoss())
        r = fm(torch.zeros(1), labels, bd)
        o = CO.full_model_forward(outs, labels, bd, weight, dict(ohem_keep=keep))
        assert torch.equal(r[0], o[0]) and torch.equal(r[2], o[2])
        assert torch.equal(r[3][0], o[3][0]) and torch.equal(r[3][1], o[3][1])
        assert all(torch.equal(a, b) for a, b in zip(r[1], o[1]))
