"""CPU: the SGD / poly-LR restatement (oracle/sgd_oracle.py) is pinned against the live torch.optim.SGD -- the arithmetic
the reference delegates to (tools/train.py:139-148) -- and against the reference's adjust_learning_rate when available."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import sgd_oracle as SO


@pytest.mark.parametrize('cfg', [dict(momentum=0.9, weight_decay=5e-4, nesterov=False),       # configs/*.yaml
                                 dict(momentum=0.9, weight_decay=5e-4, nesterov=True),
                                 dict(momentum=0.0, weight_decay=0.0, nesterov=False),
                                 dict(momentum=0.8, weight_decay=1e-2, nesterov=False, dampening=0.1)], ids=str)
def test_sgd_restatement_matches_torch(cfg):
    g = torch.Generator().manual_seed(0)
    p0 = torch.randn(1000, generator=g)
    p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.SGD([p], lr=0.01, **cfg)
    pn, buf = p0.numpy().copy(), None
    for it in range(6):
        grad = torch.randn(1000, generator=g)
        lr = SO.adjust_learning_rate(0.01, 100, it)
        opt.param_groups[0]['lr'] = lr
        p.grad = grad.clone()
        opt.step()
        pn, buf = SO.sgd_step(pn, grad.numpy(), buf, lr, first=(it == 0), **cfg)
        np.testing.assert_allclose(pn, p.detach().numpy(), rtol=2e-6, atol=1e-7)


def test_poly_lr_matches_reference():
    ref = '/root/reference'
    want = [0.01 * ((1 - i / 1000.0) ** 0.9) for i in (0, 1, 500, 999)]
    got = [SO.adjust_learning_rate(0.01, 1000, i) for i in (0, 1, 500, 999)]
    assert got == want
    from pidnet_b200.optim import adjust_learning_rate

    class Opt:
        param_groups = [dict(lr=0.0)]
    assert [adjust_learning_rate(Opt, 0.01, 1000, i) for i in (0, 1, 500, 999)] == want
    if os.path.isdir(ref + '/utils'):      # the live reference (build container only)
        import types

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
        if ref not in sys.path:
            sys.path.insert(0, ref)
        from utils.utils import adjust_learning_rate as ref_alr
        assert [ref_alr(Opt, 0.01, 1000, i) for i in (0, 1, 500, 999)] == want
