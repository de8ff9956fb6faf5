"""Drop-in for the reference `models/pidnet.py`: same `PIDNet` class signature, `get_pred_model` /
`get_seg_model` constructors, `state_dict` keys and output contract
(`[x_extra_p, x_, x_extra_d]` when `augment=True`, bare logits otherwise), but `forward` runs the
B200-native engine (hand-written sm_100a kernels behind the C ABI of include/pidnet_b200.h).

The module tree below exists only to hold parameters/buffers under the reference's names so that
upstream checkpoints load and optimizers / `.to()` / `state_dict()` behave; it contains no torch
compute.  There is no CPU or cuDNN fallback: non-CUDA inputs raise.
"""
from __future__ import annotations

import ctypes as C
import logging

import torch
import torch.nn as nn

from . import _lib

BN_MOMENTUM = 0.1  # reference: bn_mom, models/pidnet.py:12


# --------------------------------------------------------------------------- parameter tree
class _Node(nn.Module):
    """Pure container; children are registered under the reference's attribute / index names."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError('pidnet_b200 container modules hold parameters only')

    def __getitem__(self, idx):  # numeric children, like nn.Sequential
        return self._modules[str(idx)]


def _conv(cin, cout, k, bias=False, groups=1):
    return nn.Conv2d(cin, cout, k, bias=bias, groups=groups)


def _bn(ch):
    return nn.BatchNorm2d(ch, momentum=BN_MOMENTUM)


def _seq(*mods):
    """Numeric child names like nn.Sequential; `None` leaves the index unused (activation / pool slots)."""
    node = _Node()
    for i, m in enumerate(mods):
        if m is not None:
            node.add_module(str(i), m)
    return node


def _basic(cin, cout, stride):
    """BasicBlock parameters (reference model_utils.py:15-26)."""
    b = _Node()
    b.conv1, b.bn1 = _conv(cin, cout, 3), _bn(cout)
    b.conv2, b.bn2 = _conv(cout, cout, 3), _bn(cout)
    if stride != 1 or cin != cout:
        b.downsample = _seq(_conv(cin, cout, 1), _bn(cout))
    return b


def _bottle(cin, planes, stride):
    """Bottleneck parameters, expansion 2 (reference model_utils.py:51-64)."""
    b = _Node()
    b.conv1, b.bn1 = _conv(cin, planes, 1), _bn(planes)
    b.conv2, b.bn2 = _conv(planes, planes, 3), _bn(planes)
    b.conv3, b.bn3 = _conv(planes, planes * 2, 1), _bn(planes * 2)
    if stride != 1 or cin != planes * 2:
        b.downsample = _seq(_conv(cin, planes * 2, 1), _bn(planes * 2))
    return b


def _stage(kind, cin, planes, blocks, stride=1):
    exp = 2 if kind is _bottle else 1
    return _seq(kind(cin, planes, stride), *[kind(planes * exp, planes, 1) for _ in range(blocks - 1)])


def _head(cin, inter, cout):
    h = _Node()
    h.bn1, h.conv1 = _bn(cin), _conv(cin, inter, 3)
    h.bn2, h.conv2 = _bn(inter), _conv(inter, cout, 1, bias=True)
    return h


def _ppm(inp, branch, outp, deep):
    s = _Node()
    for k in range(1, 5):
        s.add_module(f'scale{k}', _seq(None, _bn(inp), None, _conv(inp, branch, 1)))
    s.scale0 = _seq(_bn(inp), None, _conv(inp, branch, 1))
    if deep:
        for k in range(1, 5):
            s.add_module(f'process{k}', _seq(_bn(branch), None, _conv(branch, branch, 3)))
    else:
        s.scale_process = _seq(_bn(branch * 4), None, _conv(branch * 4, branch * 4, 3, groups=4))
    s.compression = _seq(_bn(branch * 5), None, _conv(branch * 5, outp, 1))
    s.shortcut = _seq(_bn(inp), None, _conv(inp, outp, 1))
    return s


def _pag(cin, mid):
    p = _Node()
    p.f_x = _seq(_conv(cin, mid, 1), _bn(mid))
    p.f_y = _seq(_conv(cin, mid, 1), _bn(mid))
    return p


class PIDNet(nn.Module):
    """PIDNet(m, n, num_classes, planes, ppm_planes, head_planes, augment) -- reference pidnet.py:19."""

    def __init__(self, m=2, n=3, num_classes=19, planes=64, ppm_planes=96, head_planes=128, augment=True):
        super().__init__()
        self.augment = augment
        self._cfg = dict(m=m, n=n, num_classes=num_classes, planes=planes, ppm_planes=ppm_planes,
                         head_planes=head_planes, augment=int(bool(augment)))
        P = planes
        # I branch
        self.conv1 = _seq(_conv(3, P, 3, bias=True), _bn(P), None, _conv(P, P, 3, bias=True), _bn(P), None)
        self.layer1 = _stage(_basic, P, P, m)
        self.layer2 = _stage(_basic, P, 2 * P, m, 2)
        self.layer3 = _stage(_basic, 2 * P, 4 * P, n, 2)
        self.layer4 = _stage(_basic, 4 * P, 8 * P, n, 2)
        self.layer5 = _stage(_bottle, 8 * P, 8 * P, 2, 2)
        # P branch
        self.compression3 = _seq(_conv(4 * P, 2 * P, 1), _bn(2 * P))
        self.compression4 = _seq(_conv(8 * P, 2 * P, 1), _bn(2 * P))
        self.pag3 = _pag(2 * P, P)
        self.pag4 = _pag(2 * P, P)
        self.layer3_ = _stage(_basic, 2 * P, 2 * P, m)
        self.layer4_ = _stage(_basic, 2 * P, 2 * P, m)
        self.layer5_ = _stage(_bottle, 2 * P, 2 * P, 1)
        # D branch
        if m == 2:
            self.layer3_d = _basic(2 * P, P, 1)
            self.layer4_d = _stage(_bottle, P, P, 1)
            self.diff3 = _seq(_conv(4 * P, P, 3), _bn(P))
            self.diff4 = _seq(_conv(8 * P, 2 * P, 3), _bn(2 * P))
            self.spp = _ppm(16 * P, ppm_planes, 4 * P, deep=False)
            self.dfm = _Node()
            self.dfm.conv_p = _seq(_conv(4 * P, 4 * P, 1), _bn(4 * P))
            self.dfm.conv_i = _seq(_conv(4 * P, 4 * P, 1), _bn(4 * P))
        else:
            self.layer3_d = _basic(2 * P, 2 * P, 1)
            self.layer4_d = _basic(2 * P, 2 * P, 1)
            self.diff3 = _seq(_conv(4 * P, 2 * P, 3), _bn(2 * P))
            self.diff4 = _seq(_conv(8 * P, 2 * P, 3), _bn(2 * P))
            self.spp = _ppm(16 * P, ppm_planes, 4 * P, deep=True)
            self.dfm = _Node()
            self.dfm.conv = _seq(_bn(4 * P), None, _conv(4 * P, 4 * P, 3))
        self.layer5_d = _stage(_bottle, 2 * P, 2 * P, 1)
        # heads
        if augment:
            self.seghead_p = _head(2 * P, head_planes, num_classes)
            self.seghead_d = _head(2 * P, P, 1)
        self.final_layer = _head(4 * P, head_planes, num_classes)

        for mod in self.modules():  # reference init, pidnet.py:95-100
            if isinstance(mod, nn.Conv2d):
                nn.init.kaiming_normal_(mod.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(mod, nn.BatchNorm2d):
                nn.init.constant_(mod.weight, 1)
                nn.init.constant_(mod.bias, 0)

        # engine state (not part of the state_dict)
        self._engine = None
        self._planned = None        # (N, H, W, device index)
        self._fingerprint = None
        self._options = {}
        self.use_graph = False
        # bumped by everything that rewrites parameters / buffers behind torch's back (the training engine and FusedSGD write
        # the flat buffers from CUDA kernels, which changes neither data_ptr nor tensor._version): part of the fingerprint
        # below, so an eval forward after a training epoch re-reads the weights (reference loop: train, validate, repeat)
        self._generation = 0

    # ----------------------------------------------------------------------- engine plumbing
    def set_engine_option(self, name, value):
        """Engine switches, applied at the next plan (include/pidnet_b200.h, pidnet_set_option):
        'conv_impl': 0 tcgen05 (default) | 1 SIMT cross-check;  'lanes': 3 (default) | 1;  'use_ws': 1 | 0;
        'use_pair': 1 (CTA-pair conv kernel for Cin >= 128, default) | 0 | 2;  'ws_stages': 3 (default) | 2;
        'use_stem2': 2 (fused conv1.0 -> conv1.3 kernel, pipelined form for 32-channel stems; default) | 1 (lock-step) | 0;
        'use_pyramid': 1 (single-launch pooling pyramid, default) | 0;
        'fp32_head': 0 (default) | 1: final_layer in split-bf16 (hi + lo) arithmetic, fp32-accurate given its input."""
        self._options[name] = int(value)
        self._planned = None

    def _tensors(self):
        return [(k, v) for k, v in self.state_dict(keep_vars=True).items() if v.dtype.is_floating_point]

    def _weights_fingerprint(self):
        return (self._generation,) + tuple((v.data_ptr(), v._version) for _, v in self._tensors())

    def _ensure_engine(self):
        lib = _lib.load()
        if self._engine is None:
            h = C.c_void_p()
            cfg = _lib.Cfg(**self._cfg)
            _lib.check(lib.pidnet_create(C.byref(cfg), C.byref(h)))
            self._engine = h
        return lib

    def _sync(self, N, H, W, device):
        lib = self._ensure_engine()
        fp = self._weights_fingerprint()
        key = (N, H, W, device.index)
        if self._planned == key and self._fingerprint == fp:
            return lib
        for name, value in self._options.items():
            _lib.check(lib.pidnet_set_option(self._engine, name.encode(), value))
        for k, v in self._tensors():
            t = v.detach().to('cpu', torch.float32).contiguous()
            shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
            _lib.check(lib.pidnet_set_param(self._engine, k.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
        with torch.cuda.device(device):
            _lib.check(lib.pidnet_plan(self._engine, N, H, W, None))
        self._planned, self._fingerprint = key, fp
        return lib

    def freeze(self):
        """Skip the per-call weight-change check (weights are re-read only on shape change)."""
        self._weights_fingerprint = lambda: None  # type: ignore[assignment]
        self._fingerprint = None
        return self

    def __del__(self):
        try:
            if getattr(self, '_engine', None) is not None:
                _lib.load().pidnet_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    # ----------------------------------------------------------------------- forward
    def forward(self, x):
        if self.training:
            # train-mode forward (batch statistics + running-stat update), models/pidnet.py:136-182 under nn.Module.train().
            # Under autograd the three outputs carry a graph node whose backward is the engine's backward pass, so the
            # reference's own FullModel (utils/utils.py:39) or any custom loss can wrap this model.
            if not self.augment:
                raise NotImplementedError('pidnet_b200: the train-mode engine needs augment=True (three outputs)')
            if not x.is_cuda:
                raise RuntimeError('pidnet_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback')
            if x.dim() != 4 or x.shape[1] != 3:
                raise ValueError(f'expected input [N,3,H,W], got {tuple(x.shape)}')
            trainer = self.engine_trainer()
            params = [p for p in self.parameters()]
            if torch.is_grad_enabled() and any(p.requires_grad for p in params):
                from .train import _TrainForwardFn
                return list(_TrainForwardFn.apply(trainer, x, *params))
            return trainer.forward_train(x)
        if not x.is_cuda:
            raise RuntimeError('pidnet_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback')
        if x.dim() != 4 or x.shape[1] != 3:
            raise ValueError(f'expected input [N,3,H,W], got {tuple(x.shape)}')
        x = x.contiguous().float()
        N, _, H, W = x.shape
        lib = self._sync(N, H, W, x.device)
        ncls = self._cfg['num_classes']
        h8, w8 = H // 8, W // 8
        out = torch.empty((N, ncls, h8, w8), device=x.device, dtype=torch.float32)
        outs = [out]
        if self.augment:
            outs = [torch.empty((N, ncls, h8, w8), device=x.device, dtype=torch.float32), out,
                    torch.empty((N, 1, h8, w8), device=x.device, dtype=torch.float32)]
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            _lib.check(lib.pidnet_forward(
                self._engine, C.c_void_p(stream), C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
                C.c_void_p(outs[0].data_ptr()) if self.augment else None,
                C.c_void_p(outs[2].data_ptr()) if self.augment else None, int(self.use_graph)))
        return outs if self.augment else out

    def engine_trainer(self):
        """The ONE `EngineTrainer` of this model (it owns the flat parameter / gradient buffers the parameters point into),
        shared by `FullModel`, the train-mode forward and `FusedSGD`."""
        from .train import EngineTrainer
        if getattr(self, '_engine_trainer', None) is None:
            object.__setattr__(self, '_engine_trainer', EngineTrainer(self))
        return self._engine_trainer

    # ----------------------------------------------------------------------- camera-frame input
    IMAGENET_MEAN = (0.485, 0.456, 0.406)      # datasets/base_dataset.py:27-28, tools/custom.py:18-19 (RGB order)
    IMAGENET_STD = (0.229, 0.224, 0.225)

    def forward_u8(self, frames, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None, out_p=None, out_d=None, use_graph=None):
        """Eval forward on uint8 HWC BGR frames [N,H,W,3] (what cv2.imread returns): the reference's input_transform
        (datasets/base_dataset.py:36-44: BGR->RGB, /255, -mean, /std, HWC->CHW) is fused into the stem kernel, so the
        result equals forward(input_transform(frames)) while the host->device copy is 4x smaller."""
        if self.training:
            raise NotImplementedError('forward_u8 is an inference entry point; call .eval()')
        if not frames.is_cuda:
            raise RuntimeError('pidnet_b200 runs on CUDA (sm_100a) tensors only; there is no CPU fallback')
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[3] != 3:
            raise ValueError(f'expected uint8 frames [N,H,W,3], got {frames.dtype} {tuple(frames.shape)}')
        frames = frames.contiguous()
        N, H, W, _ = frames.shape
        lib = self._sync(N, H, W, frames.device)
        ncls = self._cfg['num_classes']
        h8, w8 = H // 8, W // 8
        mk = lambda c: torch.empty((N, c, h8, w8), device=frames.device, dtype=torch.float32)
        out = mk(ncls) if out is None else out
        if self.augment:
            out_p = mk(ncls) if out_p is None else out_p
            out_d = mk(1) if out_d is None else out_d
        m3, s3 = (C.c_double * 3)(*mean), (C.c_double * 3)(*std)
        stream = torch.cuda.current_stream(frames.device).cuda_stream
        with torch.cuda.device(frames.device):
            _lib.check(lib.pidnet_forward_u8(
                self._engine, C.c_void_p(stream), C.c_void_p(frames.data_ptr()), m3, s3, C.c_void_p(out.data_ptr()),
                C.c_void_p(out_p.data_ptr()) if self.augment else None,
                C.c_void_p(out_d.data_ptr()) if self.augment else None,
                int(self.use_graph if use_graph is None else use_graph)))
        return [out_p, out, out_d] if self.augment else out

    def segment(self, frames, mean=IMAGENET_MEAN, std=IMAGENET_STD, out=None, logits=None):
        """tools/custom.py:86-92 on the device: uint8 BGR frames -> uint8 label maps [N,H,W] (fused input transform,
        network, x8 align_corners upsample and argmax; only 3 B/pixel go up and 1 B/pixel comes back).  `out` / `logits`:
        optional preallocated label-map / logits tensors (pointer-stable calls replay the CUDA graph when use_graph is set)."""
        from .postprocess import upsample_argmax
        outs = self.forward_u8(frames, mean, std, out=logits)
        lg = outs[1] if self.augment else outs
        return upsample_argmax(lg, frames.shape[1:3], out=out)

    # ----------------------------------------------------------------------- introspection for tests / bench
    def forward_into(self, x, out, out_p=None, out_d=None, use_graph=True):
        """Pointer-stable variant (no allocation): writes into caller tensors; used by bench.py."""
        N, _, H, W = x.shape
        lib = self._sync(N, H, W, x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.pidnet_forward(
            self._engine, C.c_void_p(stream), C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
            C.c_void_p(out_p.data_ptr()) if out_p is not None else None,
            C.c_void_p(out_d.data_ptr()) if out_d is not None else None, int(use_graph)))
        return out

    def debug_tensor(self, name):
        lib = _lib.load()
        shape = (C.c_int64 * 4)()
        _lib.check(lib.pidnet_debug_tensor(self._engine, name.encode(), None, shape))
        t = torch.empty(tuple(shape), dtype=torch.float32)
        _lib.check(lib.pidnet_debug_tensor(self._engine, name.encode(), C.c_void_p(t.data_ptr()), shape))
        return t

    def profile(self, x, out, out_p=None, out_d=None):
        """Per-launch device times: list of dicts (name, kernel, lane, flops, bytes, ms), ops serialised."""
        lib = self._sync(x.shape[0], x.shape[2], x.shape[3], x.device)
        n = self.num_launches()
        ms = (C.c_float * n)()
        stream = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(lib.pidnet_profile(
            self._engine, C.c_void_p(stream), C.c_void_p(x.data_ptr()), C.c_void_p(out.data_ptr()),
            C.c_void_p(out_p.data_ptr()) if out_p is not None else None,
            C.c_void_p(out_d.data_ptr()) if out_d is not None else None, ms, n))
        rows = []
        for i in range(n):
            name, kern = C.create_string_buffer(128), C.create_string_buffer(64)
            fl, by, lane = C.c_double(), C.c_double(), C.c_int()
            _lib.check(lib.pidnet_op_info(self._engine, i, name, 128, kern, 64, C.byref(fl), C.byref(by),
                                          C.byref(lane)))
            rows.append(dict(name=name.value.decode(), kernel=kern.value.decode(), lane=lane.value, flops=fl.value,
                             bytes=by.value, ms=float(ms[i])))
        return rows

    def num_launches(self):
        return int(_lib.load().pidnet_num_launches(self._engine))

    def conv_flops(self):
        return float(_lib.load().pidnet_conv_flops(self._engine))


def _config_for(name):
    # reference dispatch (pidnet.py:186-191, 220-225): 's' in name -> S, elif 'm' in name -> M, else L
    if 's' in name:
        return dict(m=2, n=3, planes=32, ppm_planes=96, head_planes=128)
    if 'm' in name:
        return dict(m=2, n=3, planes=64, ppm_planes=96, head_planes=128)
    return dict(m=3, n=4, planes=64, ppm_planes=112, head_planes=256)


def get_pred_model(name, num_classes):
    """reference models/pidnet.py:218-227 (augment=False: forward returns the logits tensor)."""
    return PIDNet(num_classes=num_classes, augment=False, **_config_for(name))


def get_seg_model(cfg, imgnet_pretrained):
    """reference models/pidnet.py:184-216: augment=True model + checkpoint loading.

    imgnet_pretrained: file holds {'state_dict': bare keys}; else a FullModel state_dict whose keys
    carry a 6-char 'model.' prefix (optionally nested under 'state_dict').  Entries are kept when the
    key exists and the shape matches; loading is non-strict."""
    model = PIDNet(num_classes=cfg.DATASET.NUM_CLASSES, augment=True, **_config_for(cfg.MODEL.NAME))
    ckpt = torch.load(cfg.MODEL.PRETRAINED, map_location='cpu')
    own = model.state_dict()
    if imgnet_pretrained:
        src = ckpt['state_dict']
        picked = {k: v for k, v in src.items() if k in own and v.shape == own[k].shape}
    else:
        src = ckpt['state_dict'] if 'state_dict' in ckpt else ckpt
        picked = {k[6:]: v for k, v in src.items() if k[6:] in own and v.shape == own[k[6:]].shape}
    logging.info('Attention!!!')
    logging.info('Loaded {} parameters!'.format(len(picked)))
    logging.info('Over!!!')
    own.update(picked)
    model.load_state_dict(own, strict=False)
    return model
