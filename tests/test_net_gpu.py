"""Whole-network parity: the engine (through PIDNet.forward -> C ABI) vs the CPU oracle on the same
seeded weights and inputs.

Tolerances (bf16 engine, fp32 oracle): the north star asks rel. error <= 2e-2 and argmax agreement
>= 99.9 % on the logits -- asserted on trained-like weights in tests/test_trained_parity_gpu.py.
These tests use seeded synthetic weights (oracle.make_state_dict) and add a sharper LOCAL check:
every oracle stage is recomputed in fp32 from the engine's own stage inputs and must match the
engine's stage output to LOCAL_TOL = 1e-2 rel-L2 (half the north-star bf16 bound; one bf16 rounding
per stored tensor measures 2-6e-3).  End to end (error accumulated over ~60 bf16-stored tensors) the
logits must stay within E2E_TOL = 5e-2 on these untrained weights (torch's own all-bf16 forward measures
~2e-2 on S; the deeper L accumulates ~3e-2).
"""
import pytest
import torch

from oracle import pidnet_oracle as O
from pidnet_b200 import PIDNet

pytestmark = pytest.mark.gpu

LOCAL_TOL = 1e-2
E2E_TOL = 5e-2


def _dev():
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    return torch.device('cuda:0')


def build(name, ncls, augment, seed, dev, **opts):
    cfg = O.config_for(name, ncls, augment)
    sd = O.make_state_dict(cfg, seed)
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=ncls, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=augment)
    missing, unexpected = model.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    model = model.to(dev).eval()
    for k, v in opts.items():
        model.set_engine_option(k, v)
    return model, sd


def stage_report(model, sd, x, got):
    """For every oracle stage: (name, local rel-L2, end-to-end rel-L2).

    local = engine stage output vs the ORACLE stage recomputed in fp32 from the ENGINE's own stage
    inputs (isolates each fused kernel group); end-to-end = vs the oracle run from the image."""
    taps = {}
    with torch.no_grad():
        O.pidnet_forward(sd, x, taps=taps)
    outs = {'out': got[1] if isinstance(got, (list, tuple)) else got}
    if isinstance(got, (list, tuple)):
        outs['out_p'], outs['out_d'] = got[0], got[2]
    eng = {'x': x}
    rows = []
    for nm in O.stage_names(sd):
        eng[nm] = outs[nm].detach().float().cpu() if nm in outs else model.debug_tensor(nm)
        with torch.no_grad():
            loc = O.run_stage(sd, nm, [eng[i] for i in O.stage_inputs(sd, nm)], x.shape[-2:])
        rows.append((nm, O.rel_l2(eng[nm], loc), O.rel_l2(eng[nm], taps[nm])))
    return rows


CASES = [
    # name, classes, augment, N, H, W
    ('tiny_s', 5, True, 2, 64, 128),
    ('tiny_l', 5, True, 2, 128, 128),
    ('s', 19, False, 1, 256, 512),
    ('s', 19, True, 2, 128, 256),
    ('s', 19, False, 1, 200, 328),    # ragged everywhere: 50x82 stem tiles, 25x41 / 13x21 maps, 4x6 pooling pyramid input
    ('m', 11, True, 1, 360, 480),     # CamVid geometry at half size: 45x60 -> 23x30 -> 12x15 -> 6x8 (odd sizes)
    ('l', 19, True, 1, 128, 256),
]


@pytest.mark.parametrize('impl', [1, 0], ids=['simt', 'tcgen05'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: f'{c[0]}-{c[4]}x{c[5]}-b{c[3]}-aug{int(c[2])}')
def test_forward_matches_oracle(case, impl):
    dev = _dev()
    name, ncls, aug, N, H, W = case
    model, sd = build(name, ncls, aug, seed=11, dev=dev, conv_impl=impl)
    x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(5))
    with torch.no_grad():
        got = model(x.to(dev))
    torch.cuda.synchronize()
    rows = stage_report(model, sd, x, got)
    msg = ' '.join(f'{n}={l:.3g}/{e:.3g}' for n, l, e in rows)
    print(f'[{name} {H}x{W} impl={impl}] stage local/end-to-end rel-L2: {msg}')
    for n, l, e in rows:
        assert l < LOCAL_TOL, f'{n}: local rel-L2 {l:.4g} ({msg})'
        if n.startswith('out'):   # the three network outputs
            assert e < E2E_TOL, f'{n}: end-to-end rel-L2 {e:.4g} ({msg})'


def test_impls_agree_and_graph_replay():
    dev = _dev()
    m0, _ = build('s', 19, True, 3, dev, conv_impl=0)
    m1, _ = build('s', 19, True, 3, dev, conv_impl=1)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(9)).to(dev)
    with torch.no_grad():
        a, b = m0(x), m1(x)
        for u, v in zip(a, b):
            # same arithmetic, different fp32 accumulation order -> 1-ulp bf16 flips that propagate
            assert O.rel_l2(u.cpu(), v.cpu()) < E2E_TOL
        # CUDA-graph replay and single-lane execution give bit-identical results to the eager 3-lane run
        m0.use_graph = True
        c = m0(x)
        c2 = m0(x)
        m2, _ = build('s', 19, True, 3, dev, conv_impl=0, lanes=1)
        d = m2(x)
    torch.cuda.synchronize()
    for u, v, w, z in zip(a, c, c2, d):
        assert torch.equal(u, v) and torch.equal(u, w) and torch.equal(u, z)


def test_batch_independence():
    """Images are independent units: batch-of-3 output == three batch-of-1 outputs (bitwise)."""
    dev = _dev()
    model, _ = build('tiny_s', 7, False, 2, dev)
    x = torch.randn(3, 3, 128, 192, generator=torch.Generator().manual_seed(1)).to(dev)
    with torch.no_grad():
        full = model(x).clone()
        singles = torch.cat([model(x[i:i + 1]).clone() for i in range(3)])
    assert torch.equal(full, singles)


def test_state_dict_update_is_picked_up():
    dev = _dev()
    model, sd = build('tiny_s', 5, False, 4, dev)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(2)).to(dev)
    with torch.no_grad():
        a = model(x).clone()
        sd2 = O.make_state_dict(O.config_for('tiny_s', 5, False), 5)
        model.load_state_dict(sd2)
        b = model(x).clone()
        ref = O.pidnet_forward(sd2, x.cpu())
    assert not torch.equal(a, b)
    assert O.rel_l2(b.cpu(), ref) < E2E_TOL


@pytest.mark.gpu
def test_forward_u8_equals_forward_of_input_transform():
    """Row f2: uint8 BGR frames through the fused input transform == the reference's host-side input_transform
    (datasets/base_dataset.py:36-44) followed by the fp32 forward -- bit-identical logits (same fp32 op order, same
    bf16 rounding of the stem operand), incl. an odd geometry and the whole `segment` pipeline."""
    import numpy as np
    from oracle import postproc_oracle as PO
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    dev = torch.device('cuda:0')
    cfg = O.config_for('pidnet_s', 19, False)
    model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=19, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                   head_planes=cfg['head_planes'], augment=False)
    model.load_state_dict(O.make_state_dict(cfg, 1))
    model = model.to(dev).eval()
    rng = np.random.default_rng(0)
    for (N, H, W) in [(2, 128, 256), (1, 72, 120)]:
        frames = rng.integers(0, 256, (N, H, W, 3), dtype=np.uint8)
        # reference input_transform, numpy float32 exactly as written there
        img = frames.astype(np.float32)[:, :, :, ::-1]
        img = img / 255.0
        img -= np.array([0.485, 0.456, 0.406], dtype=np.float64)
        img /= np.array([0.229, 0.224, 0.225], dtype=np.float64)
        x = torch.from_numpy(np.ascontiguousarray(img.transpose(0, 3, 1, 2))).to(dev)
        want = model(x)
        got = model.forward_u8(torch.from_numpy(frames).to(dev))
        assert torch.equal(got, want)
        seg = model.segment(torch.from_numpy(frames).to(dev))
        assert np.array_equal(seg.cpu().numpy(), PO.argmax_labels(want.cpu().numpy(), H, W))


@pytest.mark.parametrize('opts', [dict(use_pair=0), dict(use_pair=2), dict(ws_stages=2), dict(use_stem2=0), dict(use_stem2=1), dict(use_pyramid=0),
                                  dict(use_ws=0), dict(use_pair=0, ws_stages=2, use_stem2=0)],
                         ids=lambda o: ','.join(f'{k}={v}' for k, v in o.items()))
@pytest.mark.parametrize('name,shape', [('pidnet_s', (2, 3, 192, 320)), ('pidnet_m', (1, 3, 128, 192))])
def test_engine_options_agree(name, shape, opts):
    """Every kernel-selection switch (CTA pairs on/off/forced, 2 or 3 staging buffers, fused or unfused stem, generic conv
    kernel only) computes the same network: logits vs the default configuration within bf16 noise (the variants differ in
    accumulation order / where the stem bias is added, not in the arithmetic they implement)."""
    dev = _dev()
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(11))
    ref_model, _ = build(name, 19, True, 5, dev)
    alt_model, _ = build(name, 19, True, 5, dev, **opts)
    with torch.no_grad():
        ref = ref_model(x.to(dev))
        alt = alt_model(x.to(dev))
    # The conv variants differ only in fp32 accumulation order (measured: bit-identical or ~1e-3).  The fused stem adds the
    # conv1.0 bias inside the GEMM (bf16 hi + lo parts) instead of after it, which flips an occasional last bf16 bit of the
    # first activation; 60 bf16-stored layers of an UNTRAINED net amplify that to ~1.2e-2 (both variants sit ~2e-2 from the
    # fp32 oracle, E2E_TOL above), so that switch gets the end-to-end bound.
    tol = E2E_TOL if 'use_stem2' in opts else 1e-2
    for a, b in zip(alt, ref):
        assert a.shape == b.shape
        assert O.rel_l2(a.float().cpu(), b.float().cpu()) < tol, (opts, O.rel_l2(a.float().cpu(), b.float().cpu()))


@pytest.mark.parametrize('case', [('s', 19, 2, 256, 512), ('l', 19, 1, 128, 256), ('m', 11, 1, 360, 480)], ids=str)
def test_fp32_head_matches_fp32_arithmetic(case):
    """north star: "fp32 logits within 1e-3".  With 'fp32_head' the final segmentation head (the logits path, final_layer,
    model_utils.py:100-112) runs in split-bf16 arithmetic (weights and the hidden tensor as hi + lo bf16 pairs, fp32
    accumulation): recomputed in fp32 from the ENGINE's own input to the head, the logits must agree to 1e-3 (measured ~1e-5),
    where the plain bf16 head sits at 2-5e-3.  The layers in front of the head stay bf16."""
    dev = _dev()
    name, ncls, N, H, W = case
    x = torch.randn(N, 3, H, W, generator=torch.Generator().manual_seed(6))
    errs = {}
    for mode in (0, 1):
        model, sd = build(name, ncls, True, seed=13, dev=dev, fp32_head=mode)
        with torch.no_grad():
            got = model(x.to(dev))
        torch.cuda.synchronize()
        dfm = model.debug_tensor('dfm')                       # relu(final_layer.bn1(dfm(...))) as the engine stored it (bf16)
        with torch.no_grad():
            loc = O.run_stage(sd, 'out', [dfm], x.shape[-2:])
        errs[mode] = (O.rel_l2(got[1].cpu(), loc), float((got[1].cpu() - loc).abs().max() / loc.abs().max()))
    print(f'[fp32 head {name} {H}x{W}] logits vs fp32 head on the same input: bf16 head rel-L2 {errs[0][0]:.3g} (max {errs[0][1]:.3g}), '
          f'split-bf16 head rel-L2 {errs[1][0]:.3g} (max {errs[1][1]:.3g})')
    assert errs[1][0] < 1e-3 and errs[1][1] < 1e-3, errs
    assert errs[1][0] < 0.1 * errs[0][0], errs
