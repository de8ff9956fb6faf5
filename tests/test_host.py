"""CPU tests of the host side: the nn.Module surface (constructor dispatch, state_dict contract,
checkpoint loading semantics of get_seg_model), the C-ABI library (loads, exports every symbol the
header declares) and the fail-loudly behaviour without a GPU."""
import ctypes
import os
import re
import types

import pytest
import torch

from oracle import pidnet_oracle as O
from pidnet_b200 import PIDNet, _lib, get_pred_model, get_seg_model
from pidnet_b200 import parallel as PAR

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_state_dict_contract_matches_reference_keys():
    """Key names, shapes and counts of SURVEY.md Appendix C (453/479 keys S,M; 519/545 L)."""
    for name, n_eval, n_aug in [('pidnet_s', 453, 479), ('pidnet_m', 453, 479), ('pidnet_l', 519, 545)]:
        for aug, n in [(False, n_eval), (True, n_aug)]:
            cfg = O.config_for(name, 19, aug)
            model = PIDNet(m=cfg['m'], n=cfg['n'], num_classes=19, planes=cfg['planes'],
                           ppm_planes=cfg['ppm_planes'], head_planes=cfg['head_planes'], augment=aug)
            sd, osd = model.state_dict(), O.make_state_dict(cfg, 0)
            assert len(sd) == n
            assert set(sd) == set(osd)
            for k in sd:
                assert sd[k].shape == osd[k].shape, k
    assert tuple(get_pred_model('pidnet_s', 19).state_dict()['spp.scale_process.2.weight'].shape) == (384, 96, 3, 3)
    assert tuple(get_pred_model('pidnet_l', 19).state_dict()['final_layer.conv2.weight'].shape) == (19, 256, 1, 1)


def test_name_dispatch():
    """'s' in name -> S, elif 'm' in name -> M, else L (reference pidnet.py:186-191, 220-225)."""
    for name, planes, m in [('pidnet_s', 32, 2), ('pidnet_small', 32, 2), ('pidnet_m', 64, 2),
                            ('pidnet_medium', 64, 2), ('pidnet_l', 64, 3), ('pidnet_large', 64, 3)]:
        mod = get_pred_model(name, 11)
        assert mod._cfg['planes'] == planes and mod._cfg['m'] == m and mod.augment is False
        assert mod._cfg['num_classes'] == 11


def test_reference_init():
    model = get_pred_model('pidnet_s', 19)
    for mod in model.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            assert torch.all(mod.weight == 1) and torch.all(mod.bias == 0) and mod.momentum == 0.1
    assert model.conv1[0].bias is not None and model.conv1[3].bias is not None      # stem convs keep bias
    assert model.final_layer.conv2.bias is not None and model.layer1[0].conv1.bias is None


def _cfg(name, ncls, path):
    return types.SimpleNamespace(MODEL=types.SimpleNamespace(NAME=name, PRETRAINED=path),
                                 DATASET=types.SimpleNamespace(NUM_CLASSES=ncls))


def test_get_seg_model_checkpoint_semantics(tmp_path):
    cfg = O.config_for('tiny_s', 19, True)
    # build "checkpoints" of the S topology
    scfg = O.config_for('pidnet_s', 19, True)
    sd = O.make_state_dict(scfg, 5)
    # (1) finetuned: FullModel state_dict, 'model.'-prefixed (+ stray loss weight), nested or not
    full = {'model.' + k: v for k, v in sd.items()}
    full['sem_loss.criterion.weight'] = torch.ones(19)
    p1 = str(tmp_path / 'ft.pt')
    torch.save(full, p1)
    m1 = get_seg_model(_cfg('pidnet_small', 19, p1), imgnet_pretrained=False)
    assert m1.augment is True
    for k, v in m1.state_dict().items():
        assert torch.equal(v, sd[k]), k
    p2 = str(tmp_path / 'ft_nested.pt')
    torch.save({'state_dict': full}, p2)
    m2 = get_seg_model(_cfg('pidnet_small', 19, p2), imgnet_pretrained=False)
    assert torch.equal(m2.state_dict()['layer3.0.conv1.weight'], sd['layer3.0.conv1.weight'])
    # (2) imagenet: bare keys under 'state_dict'; mismatching shapes (11-class head) are filtered
    img = {k: v for k, v in O.make_state_dict(O.config_for('pidnet_s', 11, True), 6).items()}
    p3 = str(tmp_path / 'imnet.pt')
    torch.save({'state_dict': img}, p3)
    m3 = get_seg_model(_cfg('pidnet_small', 19, p3), imgnet_pretrained=True)
    assert torch.equal(m3.state_dict()['conv1.0.weight'], img['conv1.0.weight'])
    assert m3.state_dict()['final_layer.conv2.weight'].shape[0] == 19          # kept its own 19-class head
    del cfg


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    with open(os.path.join(ROOT, 'include', 'pidnet_b200.h')) as f:
        header = f.read()
    declared = set(re.findall(r'\b(pidnet_[a-z0-9_]+)\s*\(', header))
    declared -= {'pidnet_engine', 'pidnet_cfg'}
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.pidnet_abi_version() == 2


def test_abi_error_paths_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.Cfg(m=5, n=3, num_classes=19, planes=32, ppm_planes=96, head_planes=128, augment=0)
    assert lib.pidnet_create(ctypes.byref(bad), ctypes.byref(h)) != 0
    assert b'm must be' in lib.pidnet_last_error()
    ok = _lib.Cfg(m=2, n=3, num_classes=19, planes=32, ppm_planes=96, head_planes=128, augment=0)
    assert lib.pidnet_create(ctypes.byref(ok), ctypes.byref(h)) == 0
    assert lib.pidnet_set_option(h, b'nonsense', 1) != 0
    assert lib.pidnet_forward(h, None, None, None, None, None, 0) != 0          # not planned
    assert b'before pidnet_plan' in lib.pidnet_last_error()
    assert lib.pidnet_destroy(h) == 0


def test_forward_fails_loudly_off_gpu():
    model = get_pred_model('pidnet_s', 19).eval()
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        model(torch.randn(1, 3, 64, 64))
    model.train()
    with pytest.raises(NotImplementedError):
        model(torch.randn(1, 3, 64, 64))


def test_product_does_not_import_oracle():
    """The product package must never route through oracle/ (or any CPU fallback)."""
    pkg = os.path.join(ROOT, 'pidnet_b200')
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(('.py', '.cu', '.cuh', '.cpp', '.h')):
                with open(os.path.join(dirpath, fn)) as f:
                    src = f.read()
                assert 'oracle' not in src.replace('no oracle', ''), f'{fn} mentions the oracle'


def test_shard_ranges_partition_the_batch():
    for n, world in [(32, 1), (32, 8), (33, 4), (5, 8), (0, 2)]:
        spans = [PAR.shard_range(n, world, r) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for (a, b), (c, d) in zip(spans, spans[1:]):
            assert b == c and a <= b and c <= d
        sizes = [b - a for a, b in spans]
        assert max(sizes) - min(sizes) <= 1


def test_fastdiv_is_exact():
    """The elementwise kernels decode (channel group, w, h, n) from the thread index with a launch-time magic-number division
    (Granlund-Montgomery round-up, kernels.cu); a wrong quotient would silently permute pixels.  Exhaustive over small n for
    every divisor the engine can meet (channel groups, widths, heights up to 4096) plus edge and random 32-bit dividends."""
    import random
    lib = _lib.load()
    rng = random.Random(0)
    edge = [0, 1, 2, 2 ** 16 - 1, 2 ** 16, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 2, 2 ** 32 - 1]
    for d in list(range(1, 4097)) + [2 ** 20 + 7, 2 ** 31 - 1, 2 ** 31, 2 ** 32 - 1]:
        ns = edge + [d - 1, d, d + 1, 2 * d - 1, 2 * d, 3 * d + 1] + [rng.randrange(2 ** 32) for _ in range(24)]
        if d <= 64:
            ns += list(range(0, 1024))
        for n in ns:
            n &= 0xFFFFFFFF
            assert lib.pidnet_debug_fastdiv(n, d) == n // d, (n, d)
