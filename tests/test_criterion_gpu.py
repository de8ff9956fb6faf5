"""Fused criterion kernel (through the C ABI / pidnet_b200.criterion) vs the oracle and the golden vectors.

Tolerances: fp32 per-pixel math with fp64 accumulation vs the reference's fp32 throughout.  OHEM keeps the
pixels with p STRICTLY below an order statistic of p; confident (saturated) softmaxes produce large groups
of pixels whose p agree to the last few ulps, so which tie group contains rank k -- and with it |K| -- moves
with the rounding of exp/log (torch CPU vs CUDA libm).  Scalars must agree to 5e-4 relative, gradients to
2e-3 of the gradient's max magnitude."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import criterion_oracle as CO
from pidnet_b200 import BondaryLoss, FullModel, OhemCrossEntropy
from pidnet_b200.criterion import FusedCriterion, upsample_align_corners

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'criterion_*.npz')))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def run_fused(outs, labels, bd, weight, keep, dev, grads=True):
    crit = FusedCriterion(OhemCrossEntropy(255, 0.9, keep, weight), BondaryLoss(20.0))
    out, g = crit([o.to(dev) for o in outs], labels.to(dev), bd.to(dev), need_grads=grads)
    torch.cuda.synchronize()
    return out.cpu(), [t.cpu() if t is not None else None for t in g]


def close(a, b, rel=5e-4):
    return abs(a - b) <= rel * max(abs(b), 1e-6)


@pytest.mark.parametrize('path', GOLDEN, ids=os.path.basename)
def test_fused_matches_reference_golden(path):
    dev = _dev()
    z = np.load(path)
    outs = [torch.from_numpy(z[k]) for k in ('x_p', 'x_m', 'x_d')]
    labels = torch.from_numpy(z['labels'].astype(np.int64))
    bd = torch.from_numpy(z['bd'].astype(np.float32))
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS) if bool(z['weights']) else None
    out, g = run_fused(outs, labels, bd, weight, int(z['min_kept']), dev)
    assert close(out[0].item(), float(z['loss'])), (out[0].item(), float(z['loss']))
    assert close(out[1].item(), float(z['loss_s']))
    assert close(out[2].item(), float(z['loss_b']))
    assert abs(out[3].item() - float(z['acc'])) < 1e-6
    for gi, k in zip(g, ('g_p', 'g_m', 'g_d')):
        ref = z[k]
        assert np.abs(gi.numpy() - ref).max() <= 2e-3 * np.abs(ref).max(), k


@pytest.mark.parametrize('case', [(12, 19, 256, 256, 21, 131072, False), (4, 19, 512, 512, 22, 131072, True),
                                  (2, 11, 360, 480, 23, 20000, False), (2, 11, 256, 384, 25, 20000, True), (3, 19, 128, 128, 24, 50, False)], ids=str)
def test_fused_matches_oracle_large(case):
    dev = _dev()
    n, ncls, h, w, seed, keep, aligned = case
    outs, labels, bd = CO.synthetic_batch(n, ncls, h, w, seed, aligned=aligned)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS) if ncls == 19 else None
    ro = [o.clone().requires_grad_(True) for o in outs]
    losses, _, acc, ll = CO.full_model_forward(ro, labels, bd, weight, dict(ohem_keep=keep))
    losses.mean().backward()
    out, g = run_fused(outs, labels, bd, weight, keep, dev)
    assert close(out[0].item(), losses.mean().item()), (out[0].item(), losses.mean().item())
    assert close(out[1].item(), ll[0].mean().item()) and close(out[2].item(), ll[1].item())
    assert abs(out[3].item() - acc.item()) < 1e-6
    # A pixel whose target probability lies within fp32 rounding of the OHEM threshold can fall on either side of it (the oracle's
    # softmax and the kernel's round differently): it moves <= (bw1 / K1 + sb / K2) * class weight of gradient into the 4 low-res
    # cells x C channels it touches.  Everything else must agree to 2e-3 of the largest gradient.
    wmax = float(weight.max()) if weight is not None else 1.0
    flip = 2.0 * (1.0 / max(float(out[10]), 1.0) + 1.0 / max(float(out[11]), 1.0)) * wmax
    for gi, r in zip(g, ro):
        d = (gi - r.grad).abs()
        tight = 2e-3 * r.grad.abs().max().item()
        assert d.max().item() <= tight + flip, (d.max().item(), tight, flip)
        assert int((d > tight).sum()) <= 2 * 4 * ncls, 'more than two borderline pixels disagree with the oracle'


def test_upsample_and_fullmodel_surface():
    dev = _dev()
    x = torch.randn(2, 19, 16, 24, generator=torch.Generator().manual_seed(0))
    ref = torch.nn.functional.interpolate(x, size=(128, 192), mode='bilinear', align_corners=True)
    got = upsample_align_corners(x.to(dev), (128, 192)).cpu()
    assert (got - ref).abs().max() < 1e-5

    class Dummy(torch.nn.Module):
        def __init__(self, o):
            super().__init__()
            self.o = o
        def forward(self, x):
            return list(self.o)
    outs, labels, bd = CO.synthetic_batch(2, 19, 64, 128, 31)
    weight = torch.tensor(CO.CITYSCAPES_CLASS_WEIGHTS)
    fm = FullModel(Dummy([o.to(dev) for o in outs]), OhemCrossEntropy(255, 0.9, 131072, weight), BondaryLoss()).eval()
    loss, ups, acc, ll = fm(torch.zeros(1, device=dev), labels.to(dev), bd.to(dev))
    r = CO.full_model_forward(outs, labels, bd, weight, {})
    assert close(loss.mean().item(), r[0].mean().item()) and close(ll[0].mean().item(), r[3][0].mean().item())
    assert close(ll[1].mean().item(), r[3][1].item()) and abs(acc.mean().item() - r[2].item()) < 1e-6
    assert len(ups) == 2 and all((a.cpu() - b).abs().max() < 1e-4 for a, b in zip(ups, r[1]))
    # loss_map=True: the reference's exact return shapes -- loss [1,N,H,W], loss_list[0] [N,H,W] (per-pixel maps that the
    # reduction='none' aux term broadcasts into, utils/criterion.py:50-60,94 / utils/utils.py:55-57)
    fm2 = FullModel(Dummy([o.to(dev) for o in outs]), OhemCrossEntropy(255, 0.9, 131072, weight), BondaryLoss(), loss_map=True).eval()
    loss2, _, acc2, ll2 = fm2(torch.zeros(1, device=dev), labels.to(dev), bd.to(dev))
    assert tuple(loss2.shape) == tuple(r[0].shape) == (1, 2, 64, 128) and tuple(ll2[0].shape) == tuple(r[3][0].shape) == (2, 64, 128)
    assert (loss2.cpu() - r[0]).abs().max() < 2e-4 * r[0].abs().max()
    assert (ll2[0].cpu() - r[3][0]).abs().max() < 2e-4 * r[3][0].abs().max()
    assert close(loss2.mean().item(), r[0].mean().item())
