"""Helpers shared by the GPU parity tests: NHWC bf16 <-> NCHW fp32 conversion and thin wrappers
over the single-op C-ABI entry points (tests call the product ONLY through the C ABI)."""
import ctypes as C

import torch

from pidnet_b200 import _lib


def to_nhwc_bf16(x):
    """fp32 NCHW -> contiguous bf16 NHWC (device)."""
    return x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def from_nhwc(x):
    """bf16 NHWC -> fp32 NCHW."""
    return x.float().permute(0, 3, 1, 2).contiguous()


def bf16r(x):
    """round an fp32 tensor to bf16 precision (keeps fp32 dtype)."""
    return x.to(torch.bfloat16).float()


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def op_conv2d(x_nhwc, w, bias, stride, groups=1, res=None, relu=False, nchw=False, impl=0):
    lib = _lib.load()
    N, H, W, Cin = x_nhwc.shape
    Cout, _, k, _ = w.shape
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    wh = w.detach().float().cpu().contiguous()
    bh = bias.detach().float().cpu().contiguous() if bias is not None else None
    if nchw:
        out = torch.full((N, Cout, Ho, Wo), float('nan'), device=x_nhwc.device, dtype=torch.float32)
    else:
        out = torch.full((N, Ho, Wo, Cout), float('nan'), device=x_nhwc.device, dtype=torch.bfloat16)
    _lib.check(lib.pidnet_op_conv2d(_stream(), _p(x_nhwc), N, H, W, Cin, _p(wh), _p(bh), Cout, k, stride, groups,
                                    _p(res), int(relu), None if nchw else _p(out), _p(out) if nchw else None, impl))
    torch.cuda.synchronize()
    return out


def op_stem(x, w, bias):
    lib = _lib.load()
    N, _, H, W = x.shape
    Cout = w.shape[0]
    out = torch.full((N, (H + 1) // 2, (W + 1) // 2, Cout), float('nan'), device=x.device, dtype=torch.bfloat16)
    wh, bh = w.detach().float().cpu().contiguous(), bias.detach().float().cpu().contiguous()
    _lib.check(lib.pidnet_op_stem(_stream(), _p(x), N, H, W, _p(wh), _p(bh), Cout, _p(out)))
    torch.cuda.synchronize()
    return out


def op_pag(x, low, relu=True):
    lib = _lib.load()
    N, H, W, Cc = x.shape
    _, h, w, _ = low.shape
    out = torch.full_like(x, float('nan'))
    _lib.check(lib.pidnet_op_pag(_stream(), _p(x), _p(low), _p(out), N, H, W, Cc, h, w, int(relu)))
    torch.cuda.synchronize()
    return out


def op_upadd(a, b, shape, s=None, t=None, relu=False):
    lib = _lib.load()
    N, H, W, Cc = shape
    h, w = (b.shape[1], b.shape[2]) if b is not None else (1, 1)
    dev = (a if a is not None else b).device
    out = torch.full((N, H, W, Cc), float('nan'), device=dev, dtype=torch.bfloat16)
    _lib.check(lib.pidnet_op_upadd(_stream(), _p(a), _p(b), _p(out), N, H, W, Cc, h, w, _p(s), _p(t), int(relu)))
    torch.cuda.synchronize()
    return out


def op_pool(x, k, stride, pad, s=None, t=None, relu=False):
    lib = _lib.load()
    N, H, W, Cc = x.shape
    oh = 1 if k == 0 else (H + 2 * pad - k) // stride + 1
    ow = 1 if k == 0 else (W + 2 * pad - k) // stride + 1
    out = torch.full((N, oh, ow, Cc), float('nan'), device=x.device, dtype=torch.bfloat16)
    _lib.check(lib.pidnet_op_pool(_stream(), _p(x), _p(out), N, H, W, Cc, k, stride, pad, _p(s), _p(t), int(relu)))
    torch.cuda.synchronize()
    return out


def op_lightbag(p, i_low, d):
    lib = _lib.load()
    N, H, W, Cc = p.shape
    out = torch.full((N, H, W, 2 * Cc), float('nan'), device=p.device, dtype=torch.bfloat16)
    _lib.check(lib.pidnet_op_lightbag(_stream(), _p(p), _p(i_low), _p(d), _p(out), N, H, W, Cc, i_low.shape[1],
                                      i_low.shape[2]))
    torch.cuda.synchronize()
    return out


def op_bag(p, i_low, d, s, t):
    lib = _lib.load()
    N, H, W, Cc = p.shape
    out = torch.full((N, H, W, Cc), float('nan'), device=p.device, dtype=torch.bfloat16)
    _lib.check(lib.pidnet_op_bag(_stream(), _p(p), _p(i_low), _p(d), _p(out), N, H, W, Cc, i_low.shape[1],
                                 i_low.shape[2], _p(s), _p(t)))
    torch.cuda.synchronize()
    return out
