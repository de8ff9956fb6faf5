"""Drop-in for the reference's criterion on the PIDNet path: `OhemCrossEntropy`, `BondaryLoss`
(utils/criterion.py:43-132) and `FullModel` (utils/utils.py:21-57), executed by ONE fused CUDA pass
(csrc/criterion.cu) instead of materialising three full-resolution logit tensors.

Differences to the reference's surface, both deliberate:
  * the values the reference reads from the global yacs config (LOSS.BALANCE_WEIGHTS, LOSS.SB_WEIGHTS,
    MODEL.ALIGN_CORNERS, TRAIN.IGNORE_LABEL) are constructor arguments whose defaults are the values of
    the shipped YAMLs;
  * by default `FullModel.forward` returns `loss` / `loss_s` as 1-element tensors whose `.mean()` equals the
    reference's `losses.mean()` / `loss_list[0].mean()` -- the only way every reference caller consumes
    them (utils/function.py:44,56-59,118).  `FullModel(..., loss_map=True)` returns the reference's exact shapes
    instead: `loss` `[1,N,H,W]` and `loss_s` `[N,H,W]` (the per-pixel maps that `reduction='none'` broadcasting
    produces, utils/criterion.py:50-60,94), rebuilt from the kernel's per-pixel aux-head CE.
Gradients w.r.t. the three logit maps are produced by the same kernel family (`FusedCriterion.backward`).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib


class OhemCrossEntropy(nn.Module):
    """Parameter holder with the reference signature (utils/criterion.py:43-54)."""

    def __init__(self, ignore_label=-1, thres=0.7, min_kept=100000, weight=None):
        super().__init__()
        self.thresh = thres
        self.min_kept = max(1, min_kept)
        self.ignore_label = ignore_label
        # same registration as nn.CrossEntropyLoss(weight=...) -> state_dict key 'criterion.weight'
        self.criterion = nn.CrossEntropyLoss(weight=weight, ignore_index=ignore_label, reduction='none')

    def forward(self, *a, **k):
        raise RuntimeError('pidnet_b200.OhemCrossEntropy is evaluated inside FullModel (fused criterion kernel)')


class BondaryLoss(nn.Module):
    def __init__(self, coeff_bce=20.0):
        super().__init__()
        self.coeff_bce = coeff_bce

    def forward(self, *a, **k):
        raise RuntimeError('pidnet_b200.BondaryLoss is evaluated inside FullModel (fused criterion kernel)')


class FusedCriterion:
    """loss / acc (and logit gradients) of FullModel.forward from the three low-res logit maps."""

    def __init__(self, sem_loss: OhemCrossEntropy, bd_loss: BondaryLoss, balance_weights=(0.4, 1.0), sb_weights=1.0,
                 bd_threshold=0.8):
        if len(balance_weights) != 2:
            raise ValueError('lengths of prediction and target are not identical!')   # criterion.py:99
        self.cfg = _lib.CriterionCfg(ignore_label=sem_loss.ignore_label, ohem_keep=sem_loss.min_kept,
                                     ohem_thres=sem_loss.thresh, bd_threshold=bd_threshold,
                                     balance_weight_aux=balance_weights[0], balance_weight_main=balance_weights[1],
                                     sb_weight=sb_weights, coeff_bce=bd_loss.coeff_bce)
        self.sem_loss = sem_loss
        self._ws = None

    def __call__(self, outputs, labels, bd_gt, need_grads=False, aux_ce_map=None):
        lib = _lib.load()
        x_p, x_m, x_d = [o.contiguous().float() for o in outputs]
        if not x_m.is_cuda:
            raise RuntimeError('pidnet_b200 criterion runs on CUDA tensors only; there is no CPU fallback')
        N, Cc, h, w = x_m.shape
        labels = labels.contiguous().long()
        bd_gt = bd_gt.contiguous().float()
        H, W = labels.shape[1], labels.shape[2]
        need = lib.pidnet_criterion_workspace_bytes(N, H, W)
        if self._ws is None or self._ws.numel() < need or self._ws.device != x_m.device:
            self._ws = torch.empty(need, dtype=torch.uint8, device=x_m.device)
        wt = self.sem_loss.criterion.weight
        wt = wt.to(x_m.device, torch.float32).contiguous() if wt is not None else None
        out = torch.empty(16, dtype=torch.float32, device=x_m.device)
        grads = [torch.empty_like(t) for t in (x_p, x_m, x_d)] if need_grads else [None, None, None]
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        stream = torch.cuda.current_stream(x_m.device).cuda_stream
        with torch.cuda.device(x_m.device):
            _lib.check(lib.pidnet_criterion(C.c_void_p(stream), p(x_p), p(x_m), p(x_d), N, Cc, h, w, p(labels), p(bd_gt),
                                            H, W, p(wt), C.byref(self.cfg), p(self._ws), self._ws.numel(), p(out),
                                            p(grads[0]), p(grads[1]), p(grads[2]), p(aux_ce_map)))
        return out, grads


def upsample_align_corners(x, size):
    """F.interpolate(x, size, mode='bilinear', align_corners=True) (utils/utils.py:44-46) on fp32 NCHW CUDA."""
    lib = _lib.load()
    x = x.contiguous().float()
    N, Cc, h, w = x.shape
    out = torch.empty((N, Cc, size[0], size[1]), dtype=torch.float32, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    with torch.cuda.device(x.device):
        _lib.check(lib.pidnet_upsample_align_corners(C.c_void_p(stream), C.c_void_p(x.data_ptr()), N * Cc, h, w,
                                                     C.c_void_p(out.data_ptr()), size[0], size[1]))
    return out


class FullModel(nn.Module):
    """reference utils/utils.py:21-57: wraps model + losses; forward(inputs, labels, bd_gt) ->
    (loss, [up(x_extra_p), up(x_)], acc, [loss_s, loss_b])."""

    def __init__(self, model, sem_loss, bd_loss, balance_weights=(0.4, 1.0), sb_weights=1.0, return_outputs=True, loss_map=False):
        super().__init__()
        self.model = model
        self.sem_loss = sem_loss
        self.bd_loss = bd_loss
        self.return_outputs = return_outputs
        self.loss_map = loss_map
        self._bw0 = float(balance_weights[0])
        self._crit = FusedCriterion(sem_loss, bd_loss, balance_weights, sb_weights)

    def _as_maps(self, loss, loss_s, aux_ce):
        """The reference's shapes: loss_s = bw0 * CE_none(x_extra_p) + bw1 * ohem  [N,H,W]; loss = (loss_s + loss_b + loss_sb)
        .unsqueeze(0)  [1,N,H,W].  The per-pixel part enters detached and mean-free, so `.mean()` is exactly the scalar the
        kernel produced and the gradient flows through that scalar."""
        pix = self._bw0 * aux_ce
        dev = (pix - pix.mean()).detach()
        return (loss.reshape(()) + dev).unsqueeze(0), loss_s.reshape(()) + dev

    def forward(self, inputs, labels, bd_gt, *args, **kwargs):
        if self.training and torch.is_grad_enabled():
            return self._train_forward(inputs, labels, bd_gt)
        outputs = self.model(inputs, *args, **kwargs)
        aux = torch.empty(labels.shape, dtype=torch.float32, device=labels.device) if self.loss_map else None
        out, _ = self._crit(outputs, labels, bd_gt, aux_ce_map=aux)
        h, w = labels.size(1), labels.size(2)
        ups = []
        if self.return_outputs:
            ups = [o if (o.size(2) == h and o.size(3) == w) else upsample_align_corners(o, (h, w)) for o in outputs[:-1]]
        if self.loss_map:
            loss, loss_s = self._as_maps(out[0:1], out[1], aux)
            return loss, ups, out[3], [loss_s, out[2]]
        return out[0:1], ups, out[3], [out[1], out[2]]

    def _train_forward(self, inputs, labels, bd_gt):
        """Train mode: one engine call does forward (BN batch statistics), criterion and backward; the returned
        loss is attached to autograd so `loss.mean().backward()` delivers the parameter gradients."""
        from .train import _TrainStepFn
        from .pidnet import PIDNet
        if not isinstance(self.model, PIDNet):
            raise TypeError('pidnet_b200.FullModel trains pidnet_b200.PIDNet models only (call .eval() for loss evaluation)')
        trainer = self.model.engine_trainer()
        params = [p for _, p in self.model.named_parameters()]
        wt = self.sem_loss.criterion.weight
        aux = torch.empty(labels.shape, dtype=torch.float32, device=labels.device) if self.loss_map else None
        res = _TrainStepFn.apply(trainer, inputs, labels, bd_gt, wt, self._crit.cfg, aux, *params)
        loss, out12, x_p, x_m, x_d = res
        h, w = labels.size(1), labels.size(2)
        ups = [upsample_align_corners(o, (h, w)) for o in (x_p, x_m)] if self.return_outputs else []
        if self.loss_map:
            loss, loss_s = self._as_maps(loss, out12[1], aux)
            return loss, ups, out12[3], [loss_s, out12[2]]
        return loss, ups, out12[3], [out12[1], out12[2]]

    @property
    def trainer(self):
        """The `EngineTrainer` behind the train-mode forward (created on first use)."""
        return self.model.engine_trainer()

    def check_valid(self, out=None):
        """The reference raises IndexError when an OHEM set has no valid pixel (utils/criterion.py:73) and faults on labels
        outside [0, C).  The fused criterion reports both without a host sync: the loss turns NaN, an empty selection
        contributes no gradient, and the trainer raises at its next step; call this to raise NOW (synchronises)."""
        if out is None:
            if getattr(self.model, '_engine_trainer', None) is not None:
                self.model._engine_trainer.poll_errors(wait=True)
            return
        o = out.detach().cpu()
        if o.numel() > 12 and float(o[12]) > 0:
            raise RuntimeError(f'pidnet_b200: {int(o[12])} label(s) are neither ignore_label nor in [0, num_classes)')
        if float(o[8]) == 0 or float(o[9]) == 0:
            raise IndexError('index -1 is out of bounds for dimension 0 with size 0')
