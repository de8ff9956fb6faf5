"""Kernel-level parity (through the C ABI) against torch fp64 on the same bf16-rounded operands.

Tolerances: the kernels accumulate in fp32 and round ONCE to bf16 on store, so against an fp64
reference of the same (bf16-exact) operands the error is one bf16 ulp of the result (2^-8 relative)
plus fp32 accumulation noise; we allow 1e-2 relative to the tensor's scale (north-star bf16 bound is
2e-2).  fp32 outputs (logits path) are held to 1e-4.
"""
import pytest
import torch
import torch.nn.functional as F

from tests import util as U

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def _check(got, ref, tol, what=''):
    assert torch.isfinite(got).all(), f'{what}: non-finite output'
    scale = ref.abs().max().item() + 1e-6
    err = (got.double() - ref.double()).abs().max().item()
    assert err <= tol * scale, f'{what}: max abs err {err:.4g} vs scale {scale:.4g} (tol {tol})'


CONV_CASES = [
    # N, H, W, Cin, Cout, k, stride, groups, res, relu, nchw
    (2, 32, 64, 64, 64, 3, 1, 1, True, True, False),
    (1, 64, 64, 32, 32, 3, 1, 1, False, True, False),
    (1, 45, 60, 32, 64, 3, 2, 1, False, True, False),     # odd sizes, stride 2 (CamVid 1/8 -> 1/16)
    (2, 23, 30, 64, 128, 1, 1, 1, True, False, False),    # ragged M tail, flattened 1x1
    (1, 23, 30, 64, 128, 1, 2, 1, False, False, False),   # 1x1 stride-2 downsample
    (2, 16, 32, 384, 384, 3, 1, 4, False, True, False),   # PAPPM grouped conv (K tail: 96 = 64 + 32)
    (1, 16, 16, 8, 16, 3, 1, 1, False, False, False),     # tiny channels (zero-filled K)
    (2, 24, 40, 128, 19, 1, 1, 1, False, False, True),    # logits: fp32 NCHW, Cout=19
    (1, 16, 32, 128, 136, 1, 1, 1, False, False, False),  # pag low-res conv width (2C+8)
    (1, 8, 16, 256, 256, 3, 1, 1, True, True, False),     # deep K (36 k-steps)
    (1, 16, 32, 256, 256, 3, 2, 1, False, True, False),   # layer5.0.conv2-like
    (3, 1, 1, 512, 96, 1, 1, 1, False, False, False),     # global-pool branch: 1x1 maps
    (1, 128, 256, 64, 1, 1, 1, 1, False, False, True),    # seghead_d: single output channel
    # 3x3 stride-1 single-source: the weight-stationary halo-patch kernel (impl 0)
    (2, 64, 128, 128, 128, 3, 1, 1, True, True, False),   # C=128: 2 chunks, BN=64 x 2 Cout tiles
    (1, 90, 120, 128, 128, 3, 1, 1, True, False, False),  # CamVid 1/8: ragged 16x8 tiles
    (2, 64, 64, 32, 32, 3, 1, 1, True, True, False),      # layer1: CK=32 (SWIZZLE_64B patches)
    (1, 32, 64, 256, 256, 3, 1, 1, False, True, False),   # layer4: 4 chunks, BN=32 x 8 Cout tiles
    (1, 64, 128, 128, 32, 3, 1, 1, False, False, False),  # diff3
    (1, 128, 256, 128, 128, 3, 1, 1, False, True, False), # final_layer.conv1 (many tiles per CTA)
    (4, 16, 32, 112, 112, 3, 1, 1, False, False, False),  # DAPPM process conv (K tail 112 = 64 + 48)
    (1, 40, 24, 64, 19, 3, 1, 1, False, False, True),     # 3x3 with fp32 NCHW output
    # CTA-pair instance (cta_group::2, BN = CK = 64): odd tile counts (the last pair's second CTA runs past the end),
    # many iterations per pair with the three rotating staging buffers, residual on and off
    (1, 48, 8, 64, 64, 3, 1, 1, True, True, False),       # 3 tiles
    (3, 80, 40, 64, 64, 3, 1, 1, True, True, False),      # 75 tiles: odd, one per CTA
    (5, 112, 136, 64, 64, 3, 1, 1, False, True, False),   # 595 tiles: odd, ~8 iterations per pair
    (6, 112, 136, 128, 64, 3, 1, 1, True, False, False),  # two chunks, residual, ~10 iterations per pair
    (1, 16, 8, 64, 64, 3, 1, 1, False, False, False),     # a single tile (no pair possible)
]


@pytest.mark.parametrize('impl', [1, 2, 3, 0], ids=['simt', 'tcgen05-generic', 'tcgen05-ws-single-cta', 'tcgen05-default'])
@pytest.mark.parametrize('case', CONV_CASES, ids=lambda c: 'x'.join(map(str, c)))
def test_conv2d(case, impl):
    dev = _dev()
    N, H, W, Cin, Cout, k, stride, groups, use_res, relu, nchw = case
    g = torch.Generator(device='cpu').manual_seed(hash(case) % (2 ** 31))
    x = U.bf16r(torch.randn(N, Cin, H, W, generator=g)).to(dev)
    w = U.bf16r(torch.randn(Cout, Cin // groups, k, k, generator=g) / (Cin // groups * k * k) ** 0.5)
    b = torch.randn(Cout, generator=g)
    Ho, Wo = (H + stride - 1) // stride, (W + stride - 1) // stride
    res = U.bf16r(torch.randn(N, Cout, Ho, Wo, generator=g)).to(dev) if use_res else None
    ref = F.conv2d(x.double(), w.to(dev).double(), b.to(dev).double(), stride, k // 2, 1, groups)
    if res is not None:
        ref = ref + res.double()
    if relu:
        ref = ref.relu()
    out = U.op_conv2d(U.to_nhwc_bf16(x), w, b, stride, groups, U.to_nhwc_bf16(res) if use_res else None, relu,
                      nchw, impl)
    if nchw:
        _check(out, ref, 1e-4, f'conv {case}')
    else:
        _check(U.from_nhwc(out), ref, 1e-2, f'conv {case}')


def test_conv2d_impls_agree_large():
    """tcgen05 vs SIMT restatement on a full-size PIDNet-S layer (64->64 3x3 @128x256)."""
    dev = _dev()
    g = torch.Generator().manual_seed(7)
    x = U.bf16r(torch.randn(1, 64, 128, 256, generator=g)).to(dev)
    w = U.bf16r(torch.randn(64, 64, 3, 3, generator=g) / 24.0)
    b = torch.randn(64, generator=g)
    xn = U.to_nhwc_bf16(x)
    a = U.op_conv2d(xn, w, b, 1, 1, xn, True, False, 0)
    g2 = U.op_conv2d(xn, w, b, 1, 1, xn, True, False, 2)
    c = U.op_conv2d(xn, w, b, 1, 1, xn, True, False, 1)
    _check(a.float(), c.float(), 1e-2, 'ws vs simt')
    _check(g2.float(), c.float(), 1e-2, 'generic vs simt')


def test_stem():
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 66, 130, generator=g).to(dev)
    w = torch.randn(32, 3, 3, 3, generator=g) / 5.0
    b = torch.randn(32, generator=g)
    ref = F.conv2d(x.double(), w.to(dev).double(), b.to(dev).double(), 2, 1).relu()
    _check(U.from_nhwc(U.op_stem(x, w, b)), ref, 1e-2, 'stem')


# (x2, x4 and x8 upsampling with C % 16 == 0 run the mixed-precision wide kernel, everything else the flat one)
@pytest.mark.parametrize('shape', [(2, 32, 64, 64, 16, 32), (1, 90, 120, 128, 23, 30), (1, 16, 16, 16, 4, 4), (1, 32, 64, 32, 4, 8),
                                   (2, 24, 40, 64, 12, 20), (1, 32, 32, 8, 16, 16), (1, 16, 32, 64, 8, 4)])
def test_pag(shape):
    dev = _dev()
    N, H, W, Cc, h, w = shape
    g = torch.Generator().manual_seed(3)
    x = U.bf16r(torch.randn(N, Cc, H, W, generator=g)).to(dev)
    low = U.bf16r(torch.randn(N, 2 * Cc + 8, h, w, generator=g) * 0.3).to(dev)
    y, z, t = low[:, :Cc].double(), low[:, Cc:2 * Cc].double(), low[:, 2 * Cc:2 * Cc + 1].double()
    up = lambda v: F.interpolate(v, size=[H, W], mode='bilinear', align_corners=False)
    gate = torch.sigmoid((x.double() * up(z)).sum(1, keepdim=True) + up(t))
    ref = ((1 - gate) * x.double() + gate * up(y)).relu()
    out = U.op_pag(U.to_nhwc_bf16(x), U.to_nhwc_bf16(low), True)
    _check(U.from_nhwc(out), ref, 1e-2, 'pag')


@pytest.mark.parametrize('shape', [(2, 32, 64, 32, 16, 32), (1, 90, 120, 64, 23, 30), (2, 12, 15, 96, 1, 1),
                                   (1, 12, 15, 96, 3, 4)])
def test_upadd(shape):
    dev = _dev()
    N, H, W, Cc, h, w = shape
    g = torch.Generator().manual_seed(4)
    a = U.bf16r(torch.randn(N, Cc, H, W, generator=g)).to(dev)
    b = U.bf16r(torch.randn(N, Cc, h, w, generator=g)).to(dev)
    s = (0.5 + torch.rand(Cc, generator=g)).to(dev)
    t = torch.randn(Cc, generator=g).to(dev)
    up = F.interpolate(b.double(), size=[H, W], mode='bilinear', align_corners=False)
    ref = ((a.double() + up) * s.double().view(1, -1, 1, 1) + t.double().view(1, -1, 1, 1)).relu()
    out = U.op_upadd(U.to_nhwc_bf16(a), U.to_nhwc_bf16(b), (N, H, W, Cc), s, t, True)
    _check(U.from_nhwc(out), ref, 1e-2, 'upadd affine relu')
    out = U.op_upadd(U.to_nhwc_bf16(a), U.to_nhwc_bf16(b), (N, H, W, Cc), None, None, False)
    _check(U.from_nhwc(out), a.double() + up, 1e-2, 'upadd plain')
    out = U.op_upadd(U.to_nhwc_bf16(a), None, (N, H, W, Cc), s, t, True)
    _check(U.from_nhwc(out), (a.double() * s.double().view(1, -1, 1, 1) + t.double().view(1, -1, 1, 1)).relu(), 1e-2,
           'affine relu')


@pytest.mark.parametrize('hw', [(16, 32), (12, 15), (1, 2)])
@pytest.mark.parametrize('ksp', [(5, 2, 2), (9, 4, 4), (17, 8, 8), (0, 1, 0)])
def test_pool(hw, ksp):
    dev = _dev()
    H, W = hw
    k, st, pd = ksp
    Cc, N = 128, 2
    g = torch.Generator().manual_seed(5)
    x = U.bf16r(torch.randn(N, Cc, H, W, generator=g)).to(dev)
    s = (0.5 + torch.rand(Cc, generator=g)).to(dev)
    t = torch.randn(Cc, generator=g).to(dev)
    pooled = F.adaptive_avg_pool2d(x.double(), (1, 1)) if k == 0 else F.avg_pool2d(x.double(), k, st, pd)
    ref = (pooled * s.double().view(1, -1, 1, 1) + t.double().view(1, -1, 1, 1)).relu()
    out = U.op_pool(U.to_nhwc_bf16(x), k, st, pd, s, t, True)
    _check(U.from_nhwc(out), ref, 1e-2, f'pool {ksp} {hw}')


def test_lightbag_and_bag():
    dev = _dev()
    N, H, W, Cc, h, w = 2, 32, 64, 128, 4, 8
    g = torch.Generator().manual_seed(6)
    p = U.bf16r(torch.randn(N, Cc, H, W, generator=g)).to(dev)
    d = U.bf16r(torch.randn(N, Cc, H, W, generator=g)).to(dev)
    il = U.bf16r(torch.randn(N, Cc, h, w, generator=g)).to(dev)
    s = (0.5 + torch.rand(Cc, generator=g)).to(dev)
    t = torch.randn(Cc, generator=g).to(dev)
    i = F.interpolate(il.double(), size=[H, W], mode='bilinear', align_corners=False)
    e = torch.sigmoid(d.double())
    ref_uv = torch.cat([(1 - e) * i + p.double(), i + e * p.double()], 1)
    out = U.op_lightbag(U.to_nhwc_bf16(p), U.to_nhwc_bf16(il), U.to_nhwc_bf16(d))
    _check(U.from_nhwc(out), ref_uv, 1e-2, 'lightbag uv')
    ref_bag = ((e * p.double() + (1 - e) * i) * s.double().view(1, -1, 1, 1) + t.double().view(1, -1, 1, 1)).relu()
    out = U.op_bag(U.to_nhwc_bf16(p), U.to_nhwc_bf16(il), U.to_nhwc_bf16(d), s, t)
    _check(U.from_nhwc(out), ref_bag, 1e-2, 'bag blend')
