"""CPU: oracle/postproc_oracle.py pinned against live torch (F.interpolate align_corners=True, argmax) and the
reference's get_confusion_matrix formula."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import postproc_oracle as PO


@pytest.mark.parametrize('shape', [(2, 19, 16, 32, 128, 256), (1, 11, 9, 12, 72, 96), (1, 3, 5, 7, 33, 50), (1, 2, 1, 1, 8, 8)], ids=str)
def test_upsample_restatement_matches_torch(shape):
    N, C, h, w, H, W = shape
    x = torch.randn(N, C, h, w, generator=torch.Generator().manual_seed(1)) * 3
    ref = torch.nn.functional.interpolate(x, size=(H, W), mode='bilinear', align_corners=True)
    got = PO.upsample_align_corners(x.numpy(), H, W)
    np.testing.assert_allclose(got, ref.numpy(), rtol=1e-5, atol=2e-6)
    # class indices: equal wherever the decision is not inside fp32 rounding noise
    am_ref, am = ref.argmax(1).numpy(), PO.argmax_labels(x.numpy(), H, W)
    top2 = ref.topk(min(2, C), dim=1).values
    margin = (top2[:, 0] - top2[:, -1]).numpy() if C > 1 else np.ones_like(am_ref, dtype=np.float32)
    assert np.array_equal(am[margin > 1e-4], am_ref[margin > 1e-4])
    assert (am != am_ref).mean() < 1e-3


def test_confusion_matrix_restatement():
    g = np.random.default_rng(0)
    C = 5
    logits = g.standard_normal((2, C, 12, 16)).astype(np.float32)
    label = g.integers(0, C, (2, 12, 16))
    label[0, :3] = 255
    cm = PO.confusion_matrix(label, logits, C, 255)
    pred = logits.argmax(1)
    want = np.zeros((C, C))
    for t, q in zip(label[label != 255], pred[label != 255]):
        want[t, q] += 1
    assert np.array_equal(cm, want)
    iou, miou = PO.mean_iou(cm)
    assert iou.shape == (C,) and 0 <= miou <= 1
    if os.path.isdir('/root/reference/utils') and hasattr(np, 'int'):     # the reference uses np.int (numpy < 1.24)
        sys.path.insert(0, '/root/reference')
        from utils.utils import get_confusion_matrix
        ref = get_confusion_matrix(torch.from_numpy(label), torch.from_numpy(logits), (12, 16), C, 255)
        assert np.array_equal(ref, cm)
