"""Post-processing of the logits on the device: the reference upsamples the [N,C,H/8,W/8]
logits to [N,C,H,W] fp32 (159 MB per 1024x2048 image), copies them to the host and takes a numpy argmax
(datasets/base_dataset.py:136-150, tools/custom.py:90-92, utils/utils.py:129-152).  Here ONE kernel reads the low-res logits
and writes the uint8 label map and/or accumulates the confusion matrix (`pidnet_postprocess`)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


_MASK_WS = {}     # (device, numel) -> uint32 scratch for the per-cell candidate masks


def _call(logits, size, pred, labels, ignore, conf, prune=True):
    if not logits.is_cuda:
        raise RuntimeError('pidnet_b200 post-processing runs on CUDA only; there is no CPU fallback')
    lib = _lib.load()
    logits = logits.contiguous().float()
    N, Cc, h, w = logits.shape
    H, W = int(size[0]), int(size[1])
    p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    ws = None
    if prune and Cc > 1:
        key = (logits.device, N * h * w)
        ws = _MASK_WS.get(key)
        if ws is None:
            ws = _MASK_WS[key] = torch.empty(N * h * w, dtype=torch.int32, device=logits.device)
    with torch.cuda.device(logits.device):
        stream = torch.cuda.current_stream(logits.device).cuda_stream
        _lib.check(lib.pidnet_postprocess(C.c_void_p(stream), p(logits), N, Cc, h, w, H, W, p(pred), p(labels), int(ignore),
                                          p(conf), p(ws)))


def upsample_argmax(logits, size, prune=True, out=None):
    """uint8 [N,H,W] == torch.argmax(F.interpolate(logits, size, mode='bilinear', align_corners=True), dim=1).
    prune=False forces the exhaustive per-pixel class loop (same result; used by the parity tests); `out` (optional) is a
    preallocated uint8 [N,H,W] device tensor to write into."""
    pred = out if out is not None else torch.empty(logits.shape[0], int(size[0]), int(size[1]), dtype=torch.uint8,
                                                   device=logits.device)
    if pred.dtype != torch.uint8 or tuple(pred.shape) != (logits.shape[0], int(size[0]), int(size[1])) or not pred.is_contiguous():
        raise ValueError('out must be a contiguous uint8 [N,H,W] tensor')
    _call(logits, size, pred, None, -1, None, prune)
    return pred


def accumulate_confusion(logits, labels, num_class, ignore=-1, out=None):
    """Adds the (label, prediction) histogram of this batch to `out` (int64 [C,C] on the device; created if None)."""
    if logits.shape[1] != num_class:
        raise ValueError('logits have %d channels, num_class is %d' % (logits.shape[1], num_class))
    if out is None:
        out = torch.zeros(num_class, num_class, dtype=torch.int64, device=logits.device)
    labels = labels.contiguous().long()
    _call(logits, labels.shape[-2:], None, labels, ignore, out)
    return out


def get_confusion_matrix(label, pred, size, num_class, ignore=-1):
    """Same signature and return value (numpy float64 [C,C]) as utils/utils.py:129-152, except that `pred` are the
    LOW-RES logits (the x8 upsample of utils/function.py:95-98 is fused in)."""
    label = label[:, :size[-2], :size[-1]]
    cm = accumulate_confusion(pred, label.to(pred.device), num_class, ignore)
    return cm.cpu().numpy().astype(np.float64)
