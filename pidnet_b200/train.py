"""Host side of the training step (csrc/train.inc): binds the torch parameters of a `pidnet_b200.PIDNet` to the
engine (one flat fp32 parameter buffer + one flat fp32 gradient buffer, parameters re-pointed at views of it),
and exposes the step as a torch.autograd.Function so that the reference loop

    losses, _, acc, loss_list = model(images, labels, bd_gts); loss = losses.mean(); loss.backward(); optimizer.step()

(utils/function.py:43-49) works unchanged.  Gradients of different ranks are summed with ONE NCCL all-reduce of
the flat gradient buffer (`allreduce_gradients`), replacing nn.DataParallel's reduce (tools/train.py:136)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class EngineTrainer:
    def __init__(self, model):
        self.model = model
        self.lib = _lib.load()
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise RuntimeError('pidnet_b200 training runs on CUDA only; there is no CPU fallback')
        self.device = dev
        params = [(k, p) for k, p in model.named_parameters()]
        bufs = [(k, b) for k, b in model.named_buffers() if b.dtype.is_floating_point]
        n_p = sum((p.numel() + 3) // 4 * 4 for _, p in params)      # every view starts 16-byte aligned
        n_b = sum(b.numel() for _, b in bufs)
        self.flat_param = torch.zeros(n_p + 256, dtype=torch.float32, device=dev)     # +pad: biases are read in tiles of 32
        self.flat_grad = torch.zeros(n_p + 256, dtype=torch.float32, device=dev)
        self.flat_buf = torch.zeros(n_b + 256, dtype=torch.float32, device=dev)
        self.grad_views = {}
        h = C.c_void_p()
        cfg = _lib.Cfg(**model._cfg)
        _lib.check(self.lib.pidnet_train_create(C.byref(cfg), C.byref(h)))
        self.h = h
        off = 0
        with torch.no_grad():
            for k, p in params:
                n = p.numel()
                view = self.flat_param[off:off + n].view(p.shape)
                view.copy_(p.data)
                p.data = view                                   # parameter now lives in the flat buffer
                g = self.flat_grad[off:off + n].view(p.shape)
                self.grad_views[k] = g
                self._bind(k, view, g)
                off += (n + 3) // 4 * 4
            off = 0
            for k, b in bufs:
                n = b.numel()
                view = self.flat_buf[off:off + n].view(b.shape)
                view.copy_(b.data)
                b.data = view
                self._bind(k, view, None)
                off += n
            # BatchNorm step counters: one flat int64 buffer, so the per-step increment is ONE launch
            nbt = [m.num_batches_tracked for m in model.modules()
                   if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None]
            self.flat_nbt = torch.zeros(max(len(nbt), 1), dtype=torch.int64, device=dev)
            for i, b in enumerate(nbt):
                self.flat_nbt[i] = b
                b.data = self.flat_nbt[i]
        self.n_param = n_p
        self.planned = None
        self.out12 = torch.zeros(12, dtype=torch.float32, device=dev)

    def _bind(self, key, t, g):
        shape = (C.c_int64 * max(t.dim(), 1))(*t.shape)
        _lib.check(self.lib.pidnet_train_bind(self.h, key.encode(), C.c_void_p(t.data_ptr()),
                                              C.c_void_p(g.data_ptr()) if g is not None else None, shape, t.dim()))

    def __del__(self):
        try:
            if getattr(self, 'h', None) is not None:
                self.lib.pidnet_train_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def step(self, x, labels, bd_gt, class_weights, crit_cfg, backward=True, want_logits=True):
        """Runs forward(train) + criterion (+ backward into self.flat_grad). Returns (out12, [x_extra_p, x_, x_extra_d])."""
        x = x.contiguous().float()
        labels = labels.contiguous().long()
        bd_gt = bd_gt.contiguous().float()
        N, _, H, W = x.shape
        with torch.cuda.device(self.device):
            if self.planned != (N, H, W):
                _lib.check(self.lib.pidnet_train_plan(self.h, N, H, W, None))
                self.planned = (N, H, W)
            ncls = self.model._cfg['num_classes']
            outs = [None, None, None]
            if want_logits:
                outs = [torch.empty(N, ncls, H // 8, W // 8, device=self.device), torch.empty(N, ncls, H // 8, W // 8, device=self.device),
                        torch.empty(N, 1, H // 8, W // 8, device=self.device)]
            p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cw = class_weights.to(self.device, torch.float32).contiguous() if class_weights is not None else None
            _lib.check(self.lib.pidnet_train_step(self.h, C.c_void_p(stream), p(x), p(labels), p(bd_gt), p(cw), C.byref(crit_cfg),
                                                  int(backward), p(self.out12), p(outs[1]), p(outs[0]), p(outs[2])))
        self.flat_nbt += 1                                      # nn.BatchNorm2d.num_batches_tracked of every layer
        return self.out12, outs

    def forward_train(self, x):
        """Train-mode forward only (batch statistics, running-stat update): [x_extra_p, x_, x_extra_d], no autograd graph."""
        x = x.contiguous().float()
        N, _, H, W = x.shape
        with torch.cuda.device(self.device):
            if self.planned != (N, H, W):
                _lib.check(self.lib.pidnet_train_plan(self.h, N, H, W, None))
                self.planned = (N, H, W)
            ncls = self.model._cfg['num_classes']
            outs = [torch.empty(N, ncls, H // 8, W // 8, device=self.device), torch.empty(N, ncls, H // 8, W // 8, device=self.device),
                    torch.empty(N, 1, H // 8, W // 8, device=self.device)]
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.pidnet_train_forward(self.h, C.c_void_p(stream), C.c_void_p(x.data_ptr()),
                                                     C.c_void_p(outs[1].data_ptr()), C.c_void_p(outs[0].data_ptr()),
                                                     C.c_void_p(outs[2].data_ptr())))
        self.flat_nbt += 1
        return outs

    def set_option(self, name, value):
        """Engine option of the training path: "use_graph" (1 = replay CUDA graphs after the first step)."""
        _lib.check(self.lib.pidnet_train_set_option(self.h, name.encode(), int(value)))

    def debug_tensor(self, name, grad=False):
        shape = (C.c_int64 * 4)()
        _lib.check(self.lib.pidnet_train_debug_tensor(self.h, name.encode(), int(grad), None, shape))
        t = torch.empty(tuple(shape), dtype=torch.float32)
        _lib.check(self.lib.pidnet_train_debug_tensor(self.h, name.encode(), int(grad), C.c_void_p(t.data_ptr()), shape))
        return t

    def allreduce_gradients(self, average=True):
        """Sum (mean) the flat gradient over all ranks with one NCCL all-reduce (reference: DataParallel's
        reduce-add + `losses.mean()` over replicas, tools/train.py:136 / utils/function.py:44)."""
        from .parallel import allreduce_flat_gradient
        allreduce_flat_gradient(self.flat_grad, self.n_param, average)


class _TrainStepFn(torch.autograd.Function):
    """loss = step(...); backward hands the already-computed parameter gradients to autograd."""

    @staticmethod
    def forward(ctx, trainer, x, labels, bd_gt, class_weights, crit_cfg, names, *params):
        out12, outs = trainer.step(x, labels, bd_gt, class_weights, crit_cfg, backward=True)
        ctx.trainer = trainer
        ctx.names = names
        ctx.params = params
        ctx.mark_non_differentiable(*[o for o in outs])
        return (out12[0:1].clone(), out12.clone(), *outs)

    @staticmethod
    def backward(ctx, g_loss, *unused):
        tr = ctx.trainer
        tr.allreduce_gradients()
        scale = g_loss.reshape(())
        params = ctx.params
        views = [tr.grad_views[k] for k in ctx.names]
        if all(p.grad is None or p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, views)):
            # zero-copy publication: every p.grad IS its slice of the flat gradient buffer (one scaling kernel instead of
            # one multiply + one accumulate per parameter); the usual `zero_grad(); backward(); step()` loop sees no difference
            tr.flat_grad.mul_(scale)
            for p, v in zip(params, views):
                p.grad = v
            return (None,) * (7 + len(params))
        grads = [v * scale for v in views]          # a foreign .grad exists: let autograd accumulate into it
        return (None, None, None, None, None, None, None, *grads)
