"""Optimizer step of the reference loop on the engine's flat buffers.

`FusedSGD` replaces `torch.optim.SGD(params, lr, momentum, weight_decay, nesterov)` as built in tools/train.py:139-148:
ONE kernel (`pidnet_sgd_step`) updates every parameter of a `pidnet_b200.PIDNet` in place in the trainer's flat fp32
buffer, reading the flat gradient the training step (and the NCCL all-reduce) produced.  It keeps a torch-like
`param_groups` list so the reference's `adjust_learning_rate(optimizer, base_lr, max_iters, cur_iters)`
(utils/utils.py:154-160) works on it unchanged; the same schedule is exported here."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def adjust_learning_rate(optimizer, base_lr, max_iters, cur_iters, power=0.9, nbb_mult=10):
    """Poly schedule, same signature and behaviour as utils/utils.py:154-160."""
    lr = base_lr * ((1 - float(cur_iters) / max_iters) ** power)
    optimizer.param_groups[0]['lr'] = lr
    if len(optimizer.param_groups) == 2:
        optimizer.param_groups[1]['lr'] = lr * nbb_mult
    return lr


class FusedSGD:
    """SGD with momentum / weight decay / optional Nesterov over the flat buffers of an `EngineTrainer`.

    `source` is a `pidnet_b200.FullModel` (its trainer is created on demand) or an `EngineTrainer`."""

    def __init__(self, source, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        from .train import EngineTrainer
        if isinstance(source, EngineTrainer):
            self.trainer = source
        else:
            self.trainer = source.trainer
        if nesterov and (momentum <= 0 or dampening != 0):
            raise ValueError('Nesterov momentum requires a momentum and zero dampening')     # torch.optim.SGD's check
        tr = self.trainer
        self.lib = tr.lib
        self.param_groups = [dict(params=[p for _, p in tr.model.named_parameters()], lr=lr, momentum=momentum,
                                  dampening=dampening, weight_decay=weight_decay, nesterov=nesterov)]
        self.momentum_buffer = torch.zeros_like(tr.flat_param)
        self.steps = 0

    def zero_grad(self, set_to_none=True):
        if not set_to_none:
            self.trainer.flat_grad.zero_()
            return
        for p in self.param_groups[0]['params']:
            p.grad = None

    @torch.no_grad()
    def step(self, grad_scale=1.0):
        """p -= lr * (momentum-filtered (grad_scale * g + wd * p)) for every parameter, one launch, on the current stream."""
        tr, g = self.trainer, self.param_groups[0]
        n = tr.flat_param.numel() // 4 * 4
        with torch.cuda.device(tr.device):
            stream = torch.cuda.current_stream(tr.device).cuda_stream
            _lib.check(self.lib.pidnet_sgd_step(C.c_void_p(stream), C.c_void_p(tr.flat_param.data_ptr()),
                                                C.c_void_p(tr.flat_grad.data_ptr()), C.c_void_p(self.momentum_buffer.data_ptr()),
                                                n, float(g['lr']), float(g['momentum']), float(g['dampening']),
                                                float(g['weight_decay']), int(bool(g['nesterov'])), int(self.steps == 0),
                                                float(grad_scale)))
        self.steps += 1
        tr.model._generation += 1       # the kernel rewrote the parameters in place: eval plans must re-read them

    def state_dict(self):
        return dict(momentum_buffer=self.momentum_buffer.clone(), steps=self.steps,
                    param_groups=[{k: v for k, v in self.param_groups[0].items() if k != 'params'}])

    def load_state_dict(self, sd):
        self.momentum_buffer.copy_(sd['momentum_buffer'])
        self.steps = int(sd['steps'])
        self.param_groups[0].update(sd['param_groups'][0])
