"""Parity at the BASELINE.json sizes (PIDNet-S / L 1024x2048 Cityscapes, PIDNet-M 720x960 CamVid).

The small-geometry tests in test_net_gpu.py never run the persistent kernels for more than a few tiles per CTA, never reach
the CTA-pair kernel's steady state and never fill the fused stem's pipeline.  Here the same checks run at full size, where the
CPU oracle still finishes in seconds for a single image, plus the size-independent properties of the path:
  * images are independent units: a batch of N equals N batches of one, bit for bit (also across CUDA-graph replay);
  * the stages before the pooling pyramid are translation-consistent: cropping the input by a multiple of the coarsest
    tile (64 px) leaves every interior activation unchanged, bit for bit, because a pixel's arithmetic does not depend on
    where its tile lies in the image;
  * post-processing: label maps equal the oracle's argmax of the upsampled logits, bit for bit.
"""
import numpy as np
import pytest
import torch

from oracle import pidnet_oracle as O
from oracle import postproc_oracle as PO
from tests.test_net_gpu import E2E_TOL, LOCAL_TOL, _dev, build, stage_report

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name,ncls,H,W', [('s', 19, 1024, 2048), ('m', 11, 720, 960)],
                         ids=['pidnet_s-cityscapes', 'pidnet_m-camvid'])
def test_fullsize_stagewise_parity(name, ncls, H, W):
    """Every stage of the eval forward at the benchmark geometry: engine vs the fp32 oracle recomputed from the engine's own
    stage inputs (LOCAL_TOL), logits end to end within E2E_TOL."""
    dev = _dev()
    model, sd = build(name, ncls, False, seed=21, dev=dev)
    x = torch.randn(1, 3, H, W, generator=torch.Generator().manual_seed(6))
    with torch.no_grad():
        got = model(x.to(dev))
    torch.cuda.synchronize()
    rows = stage_report(model, sd, x, got)
    msg = ' '.join(f'{n}={l:.3g}/{e:.3g}' for n, l, e in rows)
    print(f'[{name} {H}x{W}] stage local/end-to-end rel-L2: {msg}')
    for n, l, e in rows:
        assert l < LOCAL_TOL, f'{n}: local rel-L2 {l:.4g} ({msg})'
        if n.startswith('out'):
            assert e < E2E_TOL, f'{n}: end-to-end rel-L2 {e:.4g} ({msg})'


def test_fullsize_l_logits():
    """PIDNet-L 1024x2048 (DAPPM / Bag, the CTA-pair and two-CTA conv kernels at their benchmark shapes): logits vs oracle."""
    dev = _dev()
    model, sd = build('l', 19, False, seed=22, dev=dev)
    x = torch.randn(1, 3, 1024, 2048, generator=torch.Generator().manual_seed(7))
    with torch.no_grad():
        got = model(x.to(dev)).float().cpu()
        ref = O.pidnet_forward(sd, x)
    assert got.shape == ref.shape == (1, 19, 128, 256)
    assert O.rel_l2(got, ref) < E2E_TOL, O.rel_l2(got, ref)


def test_fullsize_batch_independence_and_graph():
    """Batch of 4 at 1024x2048 == four batches of one, bitwise, eager and replayed from the CUDA graph (persistent kernels
    walk ~7 tiles per CTA at batch 4 and ~2 at batch 1; the fused stem and the CTA pairs see different tile -> CTA maps)."""
    dev = _dev()
    model, _ = build('s', 19, False, seed=23, dev=dev)
    x = torch.randn(4, 3, 1024, 2048, generator=torch.Generator().manual_seed(8)).to(dev)
    with torch.no_grad():
        full = model(x).clone()
        singles = torch.cat([model(x[i:i + 1]).clone() for i in range(4)])
        model.use_graph = True
        replay = model(x).clone()
        replay2 = model(x).clone()
    torch.cuda.synchronize()
    assert torch.equal(full, singles)
    assert torch.equal(full, replay) and torch.equal(full, replay2)


def test_fullsize_translation_consistency():
    """conv1 / layer1 / layer2 (everything before the global pooling) of a 1024x2048 image vs its 512x1024 crop at offset
    (256, 512): identical bits wherever the receptive field stays inside the crop."""
    dev = _dev()
    model, _ = build('s', 19, False, seed=24, dev=dev)
    x = torch.randn(1, 3, 1024, 2048, generator=torch.Generator().manual_seed(9))
    oy, ox, ch, cw = 256, 512, 512, 1024
    xc = x[:, :, oy:oy + ch, ox:ox + cw].contiguous()
    with torch.no_grad():
        model(x.to(dev))
        full = {k: model.debug_tensor(k) for k in ('conv1', 'layer1', 'layer2')}
        model(xc.to(dev))
        crop = {k: model.debug_tensor(k) for k in ('conv1', 'layer1', 'layer2')}
    # (stride, receptive-field radius in input pixels, rounded up generously)
    for k, stride, rad in (('conv1', 4, 8), ('layer1', 4, 48), ('layer2', 8, 128)):
        m = (rad + stride - 1) // stride
        f = full[k][:, :, oy // stride + m:(oy + ch) // stride - m, ox // stride + m:(ox + cw) // stride - m]
        c = crop[k][:, :, m:ch // stride - m, m:cw // stride - m]
        assert f.shape == c.shape and f.numel() > 0
        assert torch.equal(f, c), (k, (f - c).abs().max().item())


def test_fullsize_label_maps_bit_exact():
    """PIDNet.segment at 1024x2048: uint8 frames in, label maps out == the oracle's x8 align_corners upsample + argmax of the
    engine's own logits, bit for bit (SURVEY 8 rows f1 / f2 at the benchmark size)."""
    dev = _dev()
    model, _ = build('s', 19, False, seed=25, dev=dev)
    frames = torch.randint(0, 256, (2, 1024, 2048, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        seg = model.segment(frames.to(dev)).cpu().numpy()
        logits = model.forward_u8(frames.to(dev)).float().cpu().numpy()
    assert seg.shape == (2, 1024, 2048) and seg.dtype == np.uint8
    assert np.array_equal(seg, PO.argmax_labels(logits, 1024, 2048))
