"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's post-processing.

Reference: `inference` datasets/base_dataset.py:136-150 (F.interpolate bilinear align_corners=True to the image size, then
`.exp()`), the argmax that follows it (tools/custom.py:92, utils/function.py:141-145) and `get_confusion_matrix`
utils/utils.py:129-152.  The interpolation arithmetic lives in torch (aten upsample_bilinear2d); it is restated here in
numpy fp32 with torch's CUDA-kernel operation order,
    h0 * (w0 * x00 + w1 * x01) + h1 * (w0 * x10 + w1 * x11),   w1 = src - floor(src), w0 = 1 - w1,  src = dst * (in-1)/(out-1)
every product and sum rounded to fp32, and pinned in tests/test_postproc_oracle.py against the live torch F.interpolate
(values to 1e-6; class indices equal wherever the top-2 margin exceeds fp32 rounding) and against the reference's own
get_confusion_matrix.  Only tests/ may import this module."""
import numpy as np


def _lerp_ac(out, inp):
    f = np.float32
    scale = f(inp - 1) / f(out - 1) if out > 1 else f(0)
    src = (scale * np.arange(out, dtype=f)).astype(f)
    i0 = np.minimum(src.astype(np.int64), inp - 1)
    i1 = i0 + (i0 < inp - 1)
    l1 = (src - i0.astype(f)).astype(f)
    return i0, i1, l1, (f(1) - l1).astype(f)


def upsample_align_corners(x, H, W):
    """x: [N,C,h,w] float32 -> [N,C,H,W] float32 (bilinear, align_corners=True)."""
    x = np.asarray(x, np.float32)
    h, w = x.shape[-2:]
    y0, y1, hy1, hy0 = _lerp_ac(H, h)
    x0, x1, wx1, wx0 = _lerp_ac(W, w)
    r0, r1 = x[:, :, y0, :], x[:, :, y1, :]
    top = (wx0 * r0[..., x0]).astype(np.float32) + (wx1 * r0[..., x1]).astype(np.float32)
    bot = (wx0 * r1[..., x0]).astype(np.float32) + (wx1 * r1[..., x1]).astype(np.float32)
    hy0, hy1 = hy0[:, None], hy1[:, None]
    return ((hy0 * top).astype(np.float32) + (hy1 * bot).astype(np.float32)).astype(np.float32)


def argmax_labels(logits, H, W):
    """uint8 [N,H,W] label map: first maximum over classes of the upsampled logits (np.argmax tie rule == torch.argmax)."""
    return np.argmax(upsample_align_corners(logits, H, W), axis=1).astype(np.uint8)


def confusion_matrix(label, pred_logits, num_class, ignore=-1):
    """utils/utils.py:129-152 on already label-sized logits: rows = ground truth, columns = prediction."""
    seg_pred = np.asarray(np.argmax(np.asarray(pred_logits).transpose(0, 2, 3, 1), axis=3), dtype=np.uint8)
    seg_gt = np.asarray(label, dtype=np.int64)
    keep = seg_gt != ignore
    seg_gt, seg_pred = seg_gt[keep], seg_pred[keep]
    index = (seg_gt * num_class + seg_pred).astype('int32')
    label_count = np.bincount(index)
    cm = np.zeros((num_class, num_class))
    for i_label in range(num_class):
        for i_pred in range(num_class):
            cur = i_label * num_class + i_pred
            if cur < len(label_count):
                cm[i_label, i_pred] = label_count[cur]
    return cm


def mean_iou(cm):
    """utils/function.py:118-124: IoU = tp / max(1, pos + res - tp)."""
    pos, res, tp = cm.sum(1), cm.sum(0), np.diag(cm)
    iou = tp / np.maximum(1.0, pos + res - tp)
    return iou, iou.mean()
