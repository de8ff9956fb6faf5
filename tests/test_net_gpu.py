"""Whole-network parity: the engine (through PIDNet.forward -> C ABI) vs the CPU oracle on the same
seeded weights and inputs.

Tolerances (bf16 engine, fp32 oracle): north star asks rel. error <= 2e-2 and argmax agreement
>= 99.9 % on trained-like weights (tests/test_trained_parity_gpu.py).  On RANDOM-INIT weights bf16
rounding alone gives ~1.5-2.5e-2 rel-L2 for ANY implementation (torch's own
all-bf16 forward measures 1.9-2.2e-2), so here the bound is 4e-2 on the logits, 3e-2 on every named
intermediate, and the tcgen05 and SIMT conv paths must agree with each other to 1e-2.
"""
import pytest
import torch

from oracle import pidnet_oracle as O
from pidnet_b200 import PIDNet

pytestmark = pytest.mark.gpu

NAMED = ['conv1', 'layer1', 'layer2', 'layer3_', 'layer3_d', 'layer3', 'pag3', 'xd3', 'layer4', 'layer4_',
         'layer4_d', 'pag4', 'xd4', 'layer5_', 'layer5_d', 'layer5', 'spp', 'dfm']


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


def report(model, taps):
    rows = []
    for nm in NAMED:
        if nm not in taps:
            continue
        got = model.debug_tensor(nm)
        ref = taps[nm]
        if nm == 'dfm':      # engine stores relu(final_layer.bn1(dfm)) -- compare through the same map
            continue
        rows.append((nm, O.rel_l2(got, ref)))
    return rows


CASES = [
    # name, classes, augment, N, H, W
    ('tiny_s', 5, True, 2, 64, 128),
    ('tiny_l', 5, True, 2, 128, 128),
    ('s', 19, False, 1, 256, 512),
    ('s', 19, True, 2, 128, 256),
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
    taps = {}
    with torch.no_grad():
        ref = O.pidnet_forward(sd, x, taps=taps)
        got = model(x.to(dev))
    torch.cuda.synchronize()
    rows = report(model, taps)
    msg = ' '.join(f'{n}={e:.3g}' for n, e in rows)
    print(f'[{name} impl={impl}] intermediates rel-L2: {msg}')
    for n, e in rows:
        assert e < 3e-2, f'{n}: rel-L2 {e:.4g} ({msg})'
    refs = ref if aug else [ref]
    gots = got if aug else [got]
    for i, (g, r) in enumerate(zip(gots, refs)):
        assert g.shape == r.shape
        e = O.rel_l2(g.cpu(), r)
        print(f'  output {i}: rel-L2 {e:.4g} argmax agreement {O.argmax_agreement(g.cpu(), r):.5f}')
        assert e < 4e-2, f'output {i}: rel-L2 {e:.4g}'


def test_impls_agree_and_graph_replay():
    dev = _dev()
    m0, _ = build('s', 19, True, 3, dev, conv_impl=0)
    m1, _ = build('s', 19, True, 3, dev, conv_impl=1)
    x = torch.randn(2, 3, 256, 256, generator=torch.Generator().manual_seed(9)).to(dev)
    with torch.no_grad():
        a, b = m0(x), m1(x)
        for u, v in zip(a, b):
            assert O.rel_l2(u.cpu(), v.cpu()) < 1e-2
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
    assert O.rel_l2(b.cpu(), ref) < 4e-2
