"""GPU: pidnet_postprocess (fused upsample + argmax + confusion matrix) against oracle/postproc_oracle.py -- index work,
so the bar is bit-exact -- and against torch on the device."""
import numpy as np
import pytest
import torch

from oracle import postproc_oracle as PO
from pidnet_b200 import postprocess as PP

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


CASES = [(2, 19, 32, 64, 256, 512),     # x8 (staged path), Cityscapes classes
         (1, 11, 90, 120, 720, 960),    # CamVid geometry
         (1, 19, 9, 12, 72, 96),        # small x8
         (1, 3, 5, 7, 33, 50),          # non-integer scale (global-load path)
         (2, 32, 4, 4, 16, 24),         # class limit
         (1, 1, 3, 3, 8, 8),            # single class
         (1, 19, 16, 16, 16, 16)]       # identity size


@pytest.mark.parametrize('shape', CASES, ids=str)
def test_upsample_argmax_bit_exact(shape):
    dev = _dev()
    N, C, h, w, H, W = shape
    g = torch.Generator().manual_seed(7)
    x = torch.randn(N, C, h, w, generator=g) * 2
    x[:, :, ::3, ::2] = x[:, :1, ::3, ::2]          # exact ties between classes: the first maximum must win
    pred = PP.upsample_argmax(x.to(dev), (H, W))
    want = PO.argmax_labels(x.numpy(), H, W)
    assert pred.dtype == torch.uint8 and tuple(pred.shape) == (N, H, W)
    assert np.array_equal(pred.cpu().numpy(), want)
    assert torch.equal(PP.upsample_argmax(x.to(dev), (H, W), prune=False), pred)      # exhaustive loop == candidate-pruned path
    # against torch on the device: identical away from fp32-rounding-level ties
    up = torch.nn.functional.interpolate(x.to(dev), size=(H, W), mode='bilinear', align_corners=True)
    assert (up.argmax(1) != pred.long()).float().mean() < (0.35 if C > 1 and h > 1 else 1.0)


@pytest.mark.parametrize('shape', CASES[:4], ids=str)
def test_confusion_matrix_bit_exact(shape):
    dev = _dev()
    N, C, h, w, H, W = shape
    g = torch.Generator().manual_seed(8)
    x = torch.randn(N, C, h, w, generator=g) * 2
    labels = torch.randint(0, C, (N, H, W), generator=g)
    labels[:, : H // 5] = 255
    cm = PP.accumulate_confusion(x.to(dev), labels.to(dev), C, 255)
    cm = PP.accumulate_confusion(x.to(dev), labels.to(dev), C, 255, out=cm)          # accumulates
    want = PO.confusion_matrix(labels.numpy(), PO.upsample_align_corners(x.numpy(), H, W), C, 255)
    assert np.array_equal(cm.cpu().numpy(), 2 * want.astype(np.int64))
    ref_like = PP.get_confusion_matrix(labels.to(dev), x.to(dev), (H, W), C, 255)
    assert ref_like.dtype == np.float64 and np.array_equal(ref_like, want)
    # size-independent property: every non-ignored pixel is counted exactly once
    assert int(cm.sum()) == 2 * int((labels != 255).sum())


def test_pruned_path_on_smooth_and_adversarial_logits():
    """The candidate pruning must be exact: smooth maps (one dominant class per cell), near-ties far below the pruning margin,
    exact ties between a low and a high class index, constant maps."""
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    N, C, h, w, H, W = 2, 19, 24, 40, 192, 320
    base = torch.nn.functional.interpolate(torch.randn(N, C, 4, 6, generator=g) * 4, size=(h, w), mode='bicubic')
    cases = [base,                                                          # smooth
             base + 1e-6 * torch.randn(N, C, h, w, generator=g),            # rounding-level perturbations
             torch.zeros(N, C, h, w),                                       # all classes tie everywhere -> class 0
             base.clone()]
    cases[3][:, 7] = cases[3][:, 2]                                         # class 7 == class 2 exactly -> 2 wins its ties
    cases.append(torch.randn(N, C, h, w, generator=g) * 1e-5)                # everything inside the margin
    for x in cases:
        got = PP.upsample_argmax(x.to(dev), (H, W))
        assert np.array_equal(got.cpu().numpy(), PO.argmax_labels(x.numpy(), H, W))
        assert torch.equal(got, PP.upsample_argmax(x.to(dev), (H, W), prune=False))


def test_all_ignored_and_errors():
    dev = _dev()
    x = torch.randn(1, 5, 4, 4).to(dev)
    labels = torch.full((1, 32, 32), 255).to(dev)
    assert int(PP.accumulate_confusion(x, labels, 5, 255).sum()) == 0
    with pytest.raises(RuntimeError):
        PP.upsample_argmax(torch.randn(1, 40, 4, 4).to(dev), (32, 32))       # > 32 classes
    with pytest.raises(RuntimeError):
        PP.upsample_argmax(torch.randn(1, 5, 4, 4), (32, 32))                # CPU tensor: no fallback
