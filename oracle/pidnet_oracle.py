"""CPU oracle for the PIDNet hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module, and only as the checker / the timed CPU baseline.  The product
path (`pidnet_b200`) never imports it and fails loudly when its CUDA library is missing.

What it is: a functional, state_dict-driven restatement of the reference forward
(`/root/reference/models/pidnet.py:136-182` + the blocks of `models/model_utils.py`) written
as the same *unfused* sequence of primitive ops the reference issues (conv -> batch_norm ->
relu ...), so that timing it on host cores is a fair stand-in for the reference's CPU forward.
The arithmetic primitives are PyTorch's (`torch.nn.functional`), exactly the third-party
dependency the reference itself delegates all arithmetic to; an
independent numpy restatement of the non-conv primitives (bilinear, avg-pool, BN) lives in
`oracle/primitives_np.py` and is checked against these in `tests/test_oracle.py`.

Pinning: the reference ships no tests / golden vectors for this path ("parity unpinned"
in-repo).  The oracle is therefore pinned against the LIVE reference imported from
`/root/reference` in the build container (`tests/test_oracle.py::test_oracle_matches_live_reference`,
skipped where the reference is absent) and against golden vectors generated from the live
reference by `tools/make_golden.py` and committed under `tests/golden/`.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5       # nn.BatchNorm2d default, model_utils.py:8
BN_MOM = 0.1        # model_utils.py:9
ALGC = False        # model_utils.py:10 / pidnet.py:13  (in-model bilinear align_corners)


# --------------------------------------------------------------------------- config
def infer_config(sd):
    """Recover the PIDNet constructor arguments (pidnet.py:19) from a state_dict."""
    P = sd['conv1.0.weight'].shape[0]
    large = 'dfm.conv.2.weight' in sd                       # Bag vs Light_Bag (pidnet.py:70,83)
    m = 3 if large else 2
    n = 0
    while f'layer3.{n}.conv1.weight' in sd:
        n += 1
    return dict(m=m, n=n, planes=P,
                ppm_planes=sd['spp.scale0.2.weight'].shape[0],
                head_planes=sd['final_layer.conv1.weight'].shape[0],
                num_classes=sd['final_layer.conv2.weight'].shape[0],
                augment='seghead_p.conv1.weight' in sd)


class _Ctx:
    """Carries the state_dict, the train/eval flag (train mode updates BN running stats in place)
    and the 1/8-resolution output size."""

    def __init__(self, sd, training=False):
        self.sd = sd
        self.training = training
        self.size8 = None


def _conv(c, x, name, stride=1, padding=0, groups=1):
    return F.conv2d(x, c.sd[name + '.weight'], c.sd.get(name + '.bias'), stride, padding, 1, groups)


def _bn(c, x, name):
    # eval: gamma*(x-mu)/sqrt(var+eps)+beta ; train: biased batch var, running stats updated
    # with momentum 0.1 and the unbiased var.
    return F.batch_norm(x, c.sd[name + '.running_mean'], c.sd[name + '.running_var'],
                        c.sd[name + '.weight'], c.sd[name + '.bias'],
                        c.training, BN_MOM, BN_EPS)


def _up(x, size):
    return F.interpolate(x, size=list(size), mode='bilinear', align_corners=ALGC)


# --------------------------------------------------------------------------- blocks
def basic_block(c, x, p, stride, no_relu):
    """model_utils.py:28-46"""
    out = F.relu(_bn(c, _conv(c, x, p + '.conv1', stride, 1), p + '.bn1'))
    out = _bn(c, _conv(c, out, p + '.conv2', 1, 1), p + '.bn2')
    residual = x
    if (p + '.downsample.0.weight') in c.sd:
        residual = _bn(c, _conv(c, x, p + '.downsample.0', stride, 0), p + '.downsample.1')
    out = out + residual
    return out if no_relu else F.relu(out)


def bottleneck(c, x, p, stride, no_relu):
    """model_utils.py:66-87 (expansion 2, stride on conv2)"""
    out = F.relu(_bn(c, _conv(c, x, p + '.conv1'), p + '.bn1'))
    out = F.relu(_bn(c, _conv(c, out, p + '.conv2', stride, 1), p + '.bn2'))
    out = _bn(c, _conv(c, out, p + '.conv3'), p + '.bn3')
    residual = x
    if (p + '.downsample.0.weight') in c.sd:
        residual = _bn(c, _conv(c, x, p + '.downsample.0', stride, 0), p + '.downsample.1')
    out = out + residual
    return out if no_relu else F.relu(out)


def make_layer(c, x, p, block, blocks, stride=1):
    """pidnet.py:103-121: block 0 keeps the block's default no_relu (BasicBlock False,
    Bottleneck True); the last block no_relu=True; the middle ones False."""
    default_no_relu = block is bottleneck
    x = block(c, x, f'{p}.0', stride, default_no_relu)
    for i in range(1, blocks):
        x = block(c, x, f'{p}.{i}', 1, i == blocks - 1)
    return x


def segmenthead(c, x, p):
    """model_utils.py:100-112 (scale_factor is always None)"""
    x = _conv(c, F.relu(_bn(c, x, p + '.bn1')), p + '.conv1', 1, 1)
    return _conv(c, F.relu(_bn(c, x, p + '.bn2')), p + '.conv2')


def _ppm_scale(c, x, p, pool):
    """scaleK = [pool] -> BN -> ReLU -> 1x1  (model_utils.py:118-142 / 200-225)"""
    i0 = 0
    if pool is not None:
        x = pool(x)
        i0 = 1
    x = F.relu(_bn(c, x, f'{p}.{i0}'))
    return _conv(c, x, f'{p}.{i0 + 2}')


_POOLS = [lambda x: F.avg_pool2d(x, 5, 2, 2),      # count_include_pad=True, ceil_mode=False
          lambda x: F.avg_pool2d(x, 9, 4, 4),
          lambda x: F.avg_pool2d(x, 17, 8, 8),
          lambda x: F.adaptive_avg_pool2d(x, (1, 1))]


def _brc(c, x, p, padding=0, groups=1):
    """BN -> ReLU -> conv sequential with indices .0 / .2"""
    return _conv(c, F.relu(_bn(c, x, p + '.0')), p + '.2', 1, padding, groups)


def pappm(c, x, p='spp'):
    """model_utils.py:247-265"""
    size = x.shape[-2:]
    x_ = _ppm_scale(c, x, p + '.scale0', None)
    scale_list = [_up(_ppm_scale(c, x, f'{p}.scale{k + 1}', _POOLS[k]), size) + x_ for k in range(4)]
    scale_out = _brc(c, torch.cat(scale_list, 1), p + '.scale_process', 1, 4)
    return _brc(c, torch.cat([x_, scale_out], 1), p + '.compression') + _brc(c, x, p + '.shortcut')


def dappm(c, x, p='spp'):
    """model_utils.py:174-194"""
    size = x.shape[-2:]
    x_list = [_ppm_scale(c, x, p + '.scale0', None)]
    for k in range(4):
        y = _up(_ppm_scale(c, x, f'{p}.scale{k + 1}', _POOLS[k]), size) + x_list[k]
        x_list.append(_brc(c, y, f'{p}.process{k + 1}', 1))
    return _brc(c, torch.cat(x_list, 1), p + '.compression') + _brc(c, x, p + '.shortcut')


def pagfm(c, x, y, p):
    """model_utils.py:292-312 with after_relu=False, with_channel=False (pidnet.py:50-51)"""
    size = x.shape[-2:]
    y_q = _up(_bn(c, _conv(c, y, p + '.f_y.0'), p + '.f_y.1'), size)
    x_k = _bn(c, _conv(c, x, p + '.f_x.0'), p + '.f_x.1')
    sim = torch.sigmoid(torch.sum(x_k * y_q, dim=1).unsqueeze(1))
    y = _up(y, size)
    return (1 - sim) * x + sim * y


def light_bag(c, p, i, d, pre='dfm'):
    """model_utils.py:328-334"""
    e = torch.sigmoid(d)
    p_add = _bn(c, _conv(c, (1 - e) * i + p, pre + '.conv_p.0'), pre + '.conv_p.1')
    i_add = _bn(c, _conv(c, i + e * p, pre + '.conv_i.0'), pre + '.conv_i.1')
    return p_add + i_add


def bag(c, p, i, d, pre='dfm'):
    """model_utils.py:375-377"""
    e = torch.sigmoid(d)
    return _brc(c, e * p + (1 - e) * i, pre + '.conv', 1)


# --------------------------------------------------------------------------- the net
def _stages(sd):
    """The forward of pidnet.py:136-182 as an ordered list of named stages
    `(name, input names, fn(ctx, *inputs))`.  Each stage's value is the single form in which the
    reference's later code observes that tensor (i.e. AFTER the in-place ReLU where there is one,
    SURVEY.md Appendix A "ReLU placement"), which is also the form the engine stores -- so the
    tests can check every stage of the engine locally, from the engine's own stage inputs."""
    cfg = infer_config(sd)
    m, n = cfg['m'], cfg['n']
    large = m == 3

    def size8(c):
        return c.size8

    def conv1(c, x):                                                      # :24-31,141
        x = F.relu(_bn(c, _conv(c, x, 'conv1.0', 2, 1), 'conv1.1'))
        return F.relu(_bn(c, _conv(c, x, 'conv1.3', 2, 1), 'conv1.4'))

    def diff_add(c, x_d, x, p):                                           # :149-152 / :161-164 (+ in-place relu :158/:169)
        return F.relu(x_d + _up(_bn(c, _conv(c, x, p + '.0', 1, 1), p + '.1'), size8(c)))

    def pag(c, x_, x, pg, comp):                                          # :148 / :160 (+ in-place relu :157/:168)
        return F.relu(pagfm(c, x_, _bn(c, _conv(c, x, comp + '.0'), comp + '.1'), pg))

    def layer4_d(c, x_d):                                                 # :60|73,158
        return basic_block(c, x_d, 'layer4_d', 1, True) if large else make_layer(c, x_d, 'layer4_d', bottleneck, 1)

    def dfm(c, p, i, d):                                                  # :170-175 + final_layer.bn1/relu (model_utils.py:102)
        i = _up(i, size8(c))
        f = bag(c, p, i, d) if large else light_bag(c, p, i, d)
        return F.relu(_bn(c, f, 'final_layer.bn1'))

    def out(c, f):                                                        # rest of segmenthead (model_utils.py:102-103)
        f = _conv(c, f, 'final_layer.conv1', 1, 1)
        return _conv(c, F.relu(_bn(c, f, 'final_layer.bn2')), 'final_layer.conv2')

    st = [
        ('conv1', ['x'], conv1),
        ('layer1', ['conv1'], lambda c, x: F.relu(make_layer(c, x, 'layer1', basic_block, m))),         # :142-143
        ('layer2', ['layer1'], lambda c, x: F.relu(make_layer(c, x, 'layer2', basic_block, m, 2))),      # :143
        ('layer3_', ['layer2'], lambda c, x: make_layer(c, x, 'layer3_', basic_block, m)),               # :144
        ('layer3_d', ['layer2'], lambda c, x: basic_block(c, x, 'layer3_d', 1, True)),                   # :145
        ('layer3', ['layer2'], lambda c, x: F.relu(make_layer(c, x, 'layer3', basic_block, n, 2))),      # :147
        ('pag3', ['layer3_', 'layer3'], lambda c, a, b: pag(c, a, b, 'pag3', 'compression3')),
        ('xd3', ['layer3_d', 'layer3'], lambda c, a, b: diff_add(c, a, b, 'diff3')),
        ('layer4', ['layer3'], lambda c, x: F.relu(make_layer(c, x, 'layer4', basic_block, n, 2))),      # :156
        ('layer4_', ['pag3'], lambda c, x: make_layer(c, x, 'layer4_', basic_block, m)),                 # :157
        ('layer4_d', ['xd3'], layer4_d),
        ('pag4', ['layer4_', 'layer4'], lambda c, a, b: pag(c, a, b, 'pag4', 'compression4')),
        ('xd4', ['layer4_d', 'layer4'], lambda c, a, b: diff_add(c, a, b, 'diff4')),
        ('layer5_', ['pag4'], lambda c, x: make_layer(c, x, 'layer5_', bottleneck, 1)),                  # :168
        ('layer5_d', ['xd4'], lambda c, x: make_layer(c, x, 'layer5_d', bottleneck, 1)),                 # :169
        ('layer5', ['layer4'], lambda c, x: make_layer(c, x, 'layer5', bottleneck, 2, 2)),               # :38,171
        ('spp', ['layer5'], (lambda c, x: dappm(c, x)) if large else (lambda c, x: pappm(c, x))),        # :69|82
        ('dfm', ['layer5_', 'spp', 'layer5_d'], dfm),
        ('out', ['dfm'], out),
    ]
    if cfg['augment']:
        # `temp_p = x_` / `temp_d = x_d` alias tensors that are ReLU-ed in place later (:154/:157,
        # :166/:169) => the aux heads see the ReLU-ed pag3 / xd4 values
        st += [('out_p', ['pag3'], lambda c, x: segmenthead(c, x, 'seghead_p')),                         # :178
               ('out_d', ['xd4'], lambda c, x: segmenthead(c, x, 'seghead_d'))]                          # :179
    return st


def stage_names(sd):
    return [s[0] for s in _stages(sd)]


def run_stage(sd, name, inputs, image_hw, training=False):
    """Recompute ONE stage from explicitly given inputs (list of tensors in the stage's input order)."""
    c = _Ctx(sd, training)
    c.size8 = (image_hw[0] // 8, image_hw[1] // 8)
    for nm, _ins, fn in _stages(sd):
        if nm == name:
            return fn(c, *inputs)
    raise KeyError(name)


def stage_inputs(sd, name):
    for nm, ins, _fn in _stages(sd):
        if nm == name:
            return list(ins)
    raise KeyError(name)


def pidnet_forward(sd, x, training=False, taps=None):
    """Restates PIDNet.forward (pidnet.py:136-182).

    Returns logits [N,C,H/8,W/8] when the state_dict has no aux heads (get_pred_model), else
    [x_extra_p, x_, x_extra_d] (get_seg_model / augment=True).  `taps` (optional dict) receives
    every named stage value."""
    c = _Ctx(sd, training)
    c.size8 = (x.shape[-2] // 8, x.shape[-1] // 8)                    # :138-139
    vals = {'x': x}
    for nm, ins, fn in _stages(sd):
        vals[nm] = fn(c, *[vals[i] for i in ins])
    if taps is not None:
        for k, v in vals.items():
            taps[k] = v.detach().clone()
    if 'out_p' in vals:                                               # :177-180
        return [vals['out_p'], vals['out'], vals['out_d']]
    return vals['out']


# --------------------------------------------------------------------------- weights
def make_state_dict(cfg, seed, randomize_bn=True, dtype=torch.float32):
    """Seeded state_dict with the reference's key names/shapes and its init
    (kaiming-normal fan_out convs, BN gamma=1 beta=0, pidnet.py:95-100); with `randomize_bn`
    the BN affine and running stats are perturbed so that folding mistakes are visible.
    Built without importing the reference so it can run on the GPU box."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    P, m, n = cfg['planes'], cfg['m'], cfg['n']
    ppm, ncls = cfg['ppm_planes'], cfg['num_classes']
    large = m == 3

    def conv(name, cin, cout, k, bias=False, groups=1):
        fan_out = cout * k * k
        sd[name + '.weight'] = torch.randn(cout, cin // groups, k, k, generator=g, dtype=dtype) * (2.0 / fan_out) ** 0.5
        if bias:
            bound = 1.0 / (cin * k * k) ** 0.5
            sd[name + '.bias'] = (torch.rand(cout, generator=g, dtype=dtype) * 2 - 1) * bound

    def bn(name, ch, gain=1.0):
        if randomize_bn:
            sd[name + '.weight'] = gain * (0.75 + 0.5 * torch.rand(ch, generator=g, dtype=dtype))
            sd[name + '.bias'] = gain * 0.1 * torch.randn(ch, generator=g, dtype=dtype)
            sd[name + '.running_mean'] = 0.1 * torch.randn(ch, generator=g, dtype=dtype)
            sd[name + '.running_var'] = 0.6 + 0.8 * torch.rand(ch, generator=g, dtype=dtype)
        else:
            sd[name + '.weight'] = torch.ones(ch, dtype=dtype)
            sd[name + '.bias'] = torch.zeros(ch, dtype=dtype)
            sd[name + '.running_mean'] = torch.zeros(ch, dtype=dtype)
            sd[name + '.running_var'] = torch.ones(ch, dtype=dtype)
        sd[name + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)

    def basic(p, cin, cout, stride):
        conv(p + '.conv1', cin, cout, 3); bn(p + '.bn1', cout)
        conv(p + '.conv2', cout, cout, 3); bn(p + '.bn2', cout)
        if stride != 1 or cin != cout:
            conv(p + '.downsample.0', cin, cout, 1); bn(p + '.downsample.1', cout)

    def bott(p, cin, planes, stride):
        conv(p + '.conv1', cin, planes, 1); bn(p + '.bn1', planes)
        conv(p + '.conv2', planes, planes, 3); bn(p + '.bn2', planes)
        conv(p + '.conv3', planes, planes * 2, 1); bn(p + '.bn3', planes * 2)
        if stride != 1 or cin != planes * 2:
            conv(p + '.downsample.0', cin, planes * 2, 1); bn(p + '.downsample.1', planes * 2)

    def layer(p, blk, cin, planes, blocks, stride=1):
        exp = 2 if blk is bott else 1
        blk(f'{p}.0', cin, planes, stride)
        for i in range(1, blocks):
            blk(f'{p}.{i}', planes * exp, planes, 1)

    def head(p, cin, inter, cout):
        bn(p + '.bn1', cin); conv(p + '.conv1', cin, inter, 3)
        bn(p + '.bn2', inter); conv(p + '.conv2', inter, cout, 1, bias=True)

    conv('conv1.0', 3, P, 3, bias=True); bn('conv1.1', P)
    conv('conv1.3', P, P, 3, bias=True); bn('conv1.4', P)
    layer('layer1', basic, P, P, m)
    layer('layer2', basic, P, 2 * P, m, 2)
    layer('layer3', basic, 2 * P, 4 * P, n, 2)
    layer('layer4', basic, 4 * P, 8 * P, n, 2)
    layer('layer5', bott, 8 * P, 8 * P, 2, 2)
    conv('compression3.0', 4 * P, 2 * P, 1); bn('compression3.1', 2 * P)
    conv('compression4.0', 8 * P, 2 * P, 1); bn('compression4.1', 2 * P)
    # Un-normalised residual stacks at random init make |x_k . y_q| ~ 100-1000, i.e. hard-saturated
    # PagFM gates that turn bf16 rounding into pixel flips for any implementation; a trained net has
    # O(1) gate logits.  The gate BNs therefore get a small gain so the synthetic weights exercise the
    # gate in its working range (sigmoid slope matters) and a tight tolerance stays meaningful.
    pag_gain = 0.05 if randomize_bn else 1.0
    for pg in ('pag3', 'pag4'):
        conv(pg + '.f_x.0', 2 * P, P, 1); bn(pg + '.f_x.1', P, pag_gain)
        conv(pg + '.f_y.0', 2 * P, P, 1); bn(pg + '.f_y.1', P, pag_gain)
    layer('layer3_', basic, 2 * P, 2 * P, m)
    layer('layer4_', basic, 2 * P, 2 * P, m)
    layer('layer5_', bott, 2 * P, 2 * P, 1)
    if not large:
        basic('layer3_d', 2 * P, P, 1)
        layer('layer4_d', bott, P, P, 1)
        conv('diff3.0', 4 * P, P, 3); bn('diff3.1', P)
    else:
        basic('layer3_d', 2 * P, 2 * P, 1)
        basic('layer4_d', 2 * P, 2 * P, 1)
        conv('diff3.0', 4 * P, 2 * P, 3); bn('diff3.1', 2 * P)
    conv('diff4.0', 8 * P, 2 * P, 3); bn('diff4.1', 2 * P)
    layer('layer5_d', bott, 2 * P, 2 * P, 1)
    # spp
    inp = 16 * P
    for k in range(5):
        i0 = 0 if k == 0 else 1
        bn(f'spp.scale{k}.{i0}', inp); conv(f'spp.scale{k}.{i0 + 2}', inp, ppm, 1)
    if large:
        for k in range(1, 5):
            bn(f'spp.process{k}.0', ppm); conv(f'spp.process{k}.2', ppm, ppm, 3)
    else:
        bn('spp.scale_process.0', ppm * 4); conv('spp.scale_process.2', ppm * 4, ppm * 4, 3, groups=4)
    bn('spp.compression.0', ppm * 5); conv('spp.compression.2', ppm * 5, 4 * P, 1)
    bn('spp.shortcut.0', inp); conv('spp.shortcut.2', inp, 4 * P, 1)
    if large:
        bn('dfm.conv.0', 4 * P); conv('dfm.conv.2', 4 * P, 4 * P, 3)
    else:
        conv('dfm.conv_p.0', 4 * P, 4 * P, 1); bn('dfm.conv_p.1', 4 * P)
        conv('dfm.conv_i.0', 4 * P, 4 * P, 1); bn('dfm.conv_i.1', 4 * P)
    if cfg.get('augment', False):
        head('seghead_p', 2 * P, cfg['head_planes'], ncls)
        head('seghead_d', 2 * P, P, 1)
    head('final_layer', 4 * P, cfg['head_planes'], ncls)
    return sd


CONFIGS = {   # pidnet.py:218-227 / :184-191
    's': dict(m=2, n=3, planes=32, ppm_planes=96, head_planes=128),
    'm': dict(m=2, n=3, planes=64, ppm_planes=96, head_planes=128),
    'l': dict(m=3, n=4, planes=64, ppm_planes=112, head_planes=256),
    # tiny variants of the same topology (any planes is legal in the reference constructor);
    # used for fast tests and travel-safe golden fixtures
    'tiny_s': dict(m=2, n=3, planes=8, ppm_planes=8, head_planes=16),
    'tiny_l': dict(m=3, n=4, planes=8, ppm_planes=8, head_planes=16),
}


def config_for(name, num_classes, augment=False):
    """Name dispatch of get_pred_model / get_seg_model (`'s' in name` -> S, elif 'm' -> M, else L;
    pidnet.py:186-191,220-225); the `tiny_*` names are ours."""
    if name in CONFIGS:
        base = CONFIGS[name]
    elif 's' in name:
        base = CONFIGS['s']
    elif 'm' in name:
        base = CONFIGS['m']
    else:
        base = CONFIGS['l']
    return dict(base, num_classes=num_classes, augment=augment)


# --------------------------------------------------------------------------- metrics
def rel_l2(a, b):
    a = a.double().flatten(); b = b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def argmax_agreement(a, b):
    return float((a.argmax(1) == b.argmax(1)).double().mean())
