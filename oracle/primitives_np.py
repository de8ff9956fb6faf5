"""Independent numpy restatement of the non-GEMM primitives the PIDNet path uses -- TEST
INFRASTRUCTURE (same rules as oracle/pidnet_oracle.py: only tests/, smoke() and bench.py's CPU
baseline may import it).

The reference delegates these to PyTorch (`F.interpolate`, `nn.AvgPool2d`, `nn.BatchNorm2d`,
`nn.Conv2d`; call sites models/model_utils.py:118-133,180-191,253-260,299-309 and
models/pidnet.py:149-152,161-164,170-173).  PyTorch publishes their definitions; the formulas below
restate them and `tests/test_oracle.py` checks them against torch, so the
CUDA kernels are validated against two independent statements of the same arithmetic.
"""
import numpy as np


def bilinear(x, out_h, out_w, align_corners=False):
    """x: [N,C,H,W].  align_corners=False: src = max(0, (dst+0.5)*in/out - 0.5); i0 = floor(src),
    i1 = min(i0+1, in-1), lam = src - i0.   align_corners=True: src = dst*(in-1)/(out-1)."""
    n, c, h, w = x.shape

    def coords(out, inn):
        d = np.arange(out, dtype=np.float32)
        if align_corners:
            src = d * (np.float32(inn - 1) / np.float32(max(out - 1, 1))) if out > 1 else np.zeros(out, np.float32)
        else:
            scale = np.float32(inn) / np.float32(out)
            src = np.maximum(scale * (d + np.float32(0.5)) - np.float32(0.5), np.float32(0))
        i0 = np.minimum(np.floor(src).astype(np.int64), inn - 1)
        i1 = np.minimum(i0 + 1, inn - 1)
        lam = (src - i0.astype(np.float32)).astype(np.float32)
        return i0, i1, lam

    h0, h1, lh = coords(out_h, h)
    w0, w1, lw = coords(out_w, w)
    lh = lh[None, None, :, None]
    lw = lw[None, None, None, :]
    top = x[:, :, h0][:, :, :, w0] * (1 - lw) + x[:, :, h0][:, :, :, w1] * lw
    bot = x[:, :, h1][:, :, :, w0] * (1 - lw) + x[:, :, h1][:, :, :, w1] * lw
    return top * (1 - lh) + bot * lh


def avg_pool(x, k, s, p):
    """nn.AvgPool2d(k, s, p) defaults: count_include_pad=True, ceil_mode=False => out = floor((in+2p-k)/s)+1,
    every window divides by k*k (zeros of the padding are counted)."""
    n, c, h, w = x.shape
    oh, ow = (h + 2 * p - k) // s + 1, (w + 2 * p - k) // s + 1
    xp = np.zeros((n, c, h + 2 * p, w + 2 * p), x.dtype)
    xp[:, :, p:p + h, p:p + w] = x
    out = np.zeros((n, c, oh, ow), x.dtype)
    for i in range(oh):
        for j in range(ow):
            out[:, :, i, j] = xp[:, :, i * s:i * s + k, j * s:j * s + k].sum(axis=(2, 3)) / (k * k)
    return out


def global_avg_pool(x):
    return x.mean(axis=(2, 3), keepdims=True)


def batch_norm_eval(x, gamma, beta, mean, var, eps=1e-5):
    """gamma*(x-mean)/sqrt(var+eps)+beta per channel."""
    sh = (1, -1, 1, 1)
    return (x - mean.reshape(sh)) / np.sqrt(var.reshape(sh) + eps) * gamma.reshape(sh) + beta.reshape(sh)


def conv2d(x, w, b=None, stride=1, pad=0, groups=1):
    """Direct convolution (small cases only): x [N,Cin,H,W], w [Cout,Cin/groups,k,k]."""
    n, cin, h, wd = x.shape
    cout, cig, k, _ = w.shape
    oh, ow = (h + 2 * pad - k) // stride + 1, (wd + 2 * pad - k) // stride + 1
    xp = np.zeros((n, cin, h + 2 * pad, wd + 2 * pad), np.float64)
    xp[:, :, pad:pad + h, pad:pad + wd] = x
    out = np.zeros((n, cout, oh, ow), np.float64)
    cog = cout // groups
    for g in range(groups):
        xs = xp[:, g * cig:(g + 1) * cig]
        ws = w[g * cog:(g + 1) * cog].astype(np.float64)
        for r in range(k):
            for q in range(k):
                patch = xs[:, :, r:r + stride * oh:stride, q:q + stride * ow:stride]
                out[:, g * cog:(g + 1) * cog] += np.einsum('nchw,oc->nohw', patch, ws[:, :, r, q])
    if b is not None:
        out += b.reshape(1, -1, 1, 1)
    return out
