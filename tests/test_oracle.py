"""CPU tests that PIN the oracle: against the live reference (build container only), against the
golden vectors generated from the live reference (tools/make_golden.py, committed under
tests/golden/), and against an independent numpy restatement of the non-GEMM primitives."""
import glob
import hashlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import pidnet_oracle as O
from oracle import primitives_np as NP

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, 'tests', 'golden', 'pidnet_*.npz')))
REF_DIR = '/root/reference'


def _sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(sd[k].detach().cpu().numpy().tobytes())
    return h.hexdigest()


def load_golden(path):
    z = np.load(path)
    name, ncls, aug = str(z['name']), int(z['num_classes']), bool(z['augment'])
    cfg = O.config_for(name, ncls, aug)
    stored = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith('w::')}
    if stored:
        sd = stored
    else:
        sd = O.make_state_dict(cfg, int(z['wseed']))
        if _sd_digest(sd) != str(z['sd_sha256']):
            pytest.skip('torch RNG stream differs from the one the golden weights were generated with')
    x = torch.from_numpy(z['x'])
    outs = [torch.from_numpy(z[f'out{i}']) for i in range(int(z['n_out']))]
    return cfg, sd, x, outs


def test_golden_vectors_exist():
    assert len(GOLDEN) >= 6


@pytest.mark.parametrize('path', GOLDEN, ids=lambda p: os.path.basename(p))
def test_oracle_matches_golden(path):
    """Golden outputs were produced by the UNMODIFIED reference model; the oracle must reproduce them."""
    cfg, sd, x, outs = load_golden(path)
    with torch.no_grad():
        got = O.pidnet_forward(sd, x)
    got = got if isinstance(got, list) else [got]
    assert len(got) == len(outs)
    for g, r in zip(got, outs):
        assert g.shape == r.shape
        assert O.rel_l2(g, r) < 1e-5, O.rel_l2(g, r)


def test_stored_weights_match_seeded_generator():
    """The tiny_s fixture stores its weights; they must equal make_state_dict(seed) bit for bit."""
    path = [p for p in GOLDEN if 'tiny_s' in p][0]
    z = np.load(path)
    sd = O.make_state_dict(O.config_for('tiny_s', int(z['num_classes']), True), int(z['wseed']))
    if _sd_digest(sd) != str(z['sd_sha256']):
        pytest.skip('torch RNG stream differs')
    for k in sd:
        assert np.array_equal(z['w::' + k], sd[k].numpy()), k


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF_DIR, 'models')), reason='reference not mounted')
@pytest.mark.parametrize('case', [('tiny_s', True, 64, 128), ('tiny_l', True, 128, 128), ('s', False, 128, 256),
                                  ('m', True, 120, 160), ('l', True, 128, 192)], ids=str)
def test_oracle_matches_live_reference(case):
    """Bit-level agreement with the reference's own nn.Module (same weights, same input), plus the
    state_dict key/shape contract."""
    sys.dont_write_bytecode = True
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import models.pidnet as REF
    name, aug, H, W = case
    cfg = O.config_for(name, 7, aug)
    ref = REF.PIDNet(m=cfg['m'], n=cfg['n'], num_classes=7, planes=cfg['planes'], ppm_planes=cfg['ppm_planes'],
                     head_planes=cfg['head_planes'], augment=aug).eval()
    sd = O.make_state_dict(cfg, seed=1)
    rsd = ref.state_dict()
    assert set(rsd) == set(sd)
    for k in rsd:
        assert rsd[k].shape == sd[k].shape, k
    ref.load_state_dict(sd)
    x = torch.randn(2, 3, H, W, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        r, o = ref(x), O.pidnet_forward(sd, x)
    r, o = (r, o) if aug else ([r], [o])
    for a, b in zip(r, o):
        assert torch.equal(a, b)
    assert O.infer_config(sd) == cfg


def test_stage_decomposition_is_consistent():
    """run_stage(...) on the taps of a full forward reproduces every tap (the local-parity tests rely on it)."""
    cfg = O.config_for('tiny_s', 5, True)
    sd = O.make_state_dict(cfg, 2)
    x = torch.randn(1, 3, 64, 64, generator=torch.Generator().manual_seed(1))
    taps = {}
    with torch.no_grad():
        O.pidnet_forward(sd, x, taps=taps)
        for nm in O.stage_names(sd):
            got = O.run_stage(sd, nm, [taps[i] for i in O.stage_inputs(sd, nm)], x.shape[-2:])
            assert torch.equal(got, taps[nm]), nm


# ---------------------------------------------------------------- numpy restatement of the primitives
@pytest.mark.parametrize('shape', [((2, 3, 16, 32), (128, 256)), ((1, 4, 23, 30), (90, 120)), ((1, 2, 1, 1), (12, 15)),
                                   ((1, 2, 6, 8), (12, 15))])
@pytest.mark.parametrize('ac', [False, True])
def test_bilinear_np(shape, ac):
    (n, c, h, w), (oh, ow) = shape
    x = torch.randn(n, c, h, w, generator=torch.Generator().manual_seed(0))
    ref = F.interpolate(x, size=[oh, ow], mode='bilinear', align_corners=ac).numpy()
    got = NP.bilinear(x.numpy(), oh, ow, ac)
    assert np.abs(got - ref).max() < 2e-5


@pytest.mark.parametrize('hw', [(16, 32), (12, 15), (1, 2), (23, 30)])
@pytest.mark.parametrize('ksp', [(5, 2, 2), (9, 4, 4), (17, 8, 8)])
def test_avgpool_np(hw, ksp):
    x = torch.randn(2, 3, *hw, generator=torch.Generator().manual_seed(0))
    ref = F.avg_pool2d(x, *ksp).numpy()
    got = NP.avg_pool(x.numpy(), *ksp)
    assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-5


def test_bn_and_conv_np():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(2, 8, 9, 11, generator=g)
    gamma, beta, mean = torch.randn(8, generator=g), torch.randn(8, generator=g), torch.randn(8, generator=g)
    var = torch.rand(8, generator=g) + 0.5
    ref = F.batch_norm(x, mean, var, gamma, beta, False, 0.1, 1e-5).numpy()
    got = NP.batch_norm_eval(x.numpy(), gamma.numpy(), beta.numpy(), mean.numpy(), var.numpy())
    assert np.abs(got - ref).max() < 1e-5
    for (k, s, p, grp) in [(3, 1, 1, 1), (3, 2, 1, 1), (1, 2, 0, 1), (3, 1, 1, 4)]:
        w = torch.randn(12, 8 // grp, k, k, generator=g)
        b = torch.randn(12, generator=g)
        ref = F.conv2d(x, w, b, s, p, 1, grp).numpy()
        got = NP.conv2d(x.numpy(), w.numpy(), b.numpy(), s, p, grp)
        assert got.shape == ref.shape and np.abs(got - ref).max() < 1e-4
