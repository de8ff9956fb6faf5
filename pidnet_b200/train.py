"""Host side of the training step (csrc/train.inc): binds the torch parameters of a `pidnet_b200.PIDNet` to the
engine (one flat fp32 parameter buffer + one flat fp32 gradient buffer, parameters re-pointed at views of it),
and exposes it to autograd in the two forms the reference uses:

  * `FullModel(model, sem_loss, bd_loss)(images, labels, bd_gts)` -> `losses.mean().backward()` (utils/function.py:43-49):
    `_TrainStepFn` -- the forward runs the train-mode network and the fused criterion, `loss.backward()` runs the
    network backward;
  * `outputs = model(inputs)` in train mode under autograd (utils/utils.py:39, any custom loss on the three outputs):
    `_TrainForwardFn` -- the backward takes the caller's logit gradients.

Gradients of different ranks are averaged with NCCL all-reduces of ranges of the flat gradient buffer, started as soon
as a range of the backward has finalised them (buckets in reverse layer order, overlapped with the rest of the
backward), replacing nn.DataParallel's reduce (tools/train.py:136)."""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


def _world():
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


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
        self._bound = []            # (tensor holder, bound view) pairs: the engine keeps raw pointers into the views
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
                self._bound.append((k, p, view.data_ptr()))
                off += (n + 3) // 4 * 4
            off = 0
            for k, b in bufs:
                n = b.numel()
                view = self.flat_buf[off:off + n].view(b.shape)
                view.copy_(b.data)
                b.data = view
                self._bind(k, view, None)
                self._bound.append((k, b, view.data_ptr()))
                off += n
            # BatchNorm step counters: one flat int64 buffer, so the per-step increment is ONE launch
            nbt = [m.num_batches_tracked for m in model.modules()
                   if isinstance(m, torch.nn.BatchNorm2d) and m.num_batches_tracked is not None]
            self.flat_nbt = torch.zeros(max(len(nbt), 1), dtype=torch.int64, device=dev)
            for i, b in enumerate(nbt):
                self.flat_nbt[i] = b
                b.data = self.flat_nbt[i]
        self._params = [p for _, p in params]
        self._views = [self.grad_views[k] for k, _ in params]
        self.n_param = n_p
        self.planned = None
        self.out16 = torch.zeros(16, dtype=torch.float32, device=dev)
        self.seq = 0                    # train-mode forwards so far: a backward must belong to the latest one
        # False (default): ONE all-reduce of the flat gradient after the backward (30.9 MB: 0.25 ms exposed at 8 GPUs over NVSwitch).
        # True: bucketed all-reduces behind the backward ranges.  Measured at 8 x B200 (profiles/r02/final/scale_n8):
        # 11.87 ms/step bucketed vs 11.41 ms single vs 11.16 ms without exchange -- the 26 NCCL kernels compete for SMs with
        # backward kernels that are sized to own the whole GPU (persistent, one CTA per SM), which costs more than the overlap hides.
        self.overlap_allreduce = False
        self._seg_ranges = None
        self._pending = None            # (event, pinned copy of out16) of the last criterion call: deferred error report
        self._host16 = torch.zeros(16, dtype=torch.float32).pin_memory()
        # nn.DataParallel replicates GPU 0's parameters and buffers every step (tools/train.py:136); one process per GPU
        # starts from rank 0's copy instead
        if _world() > 1:
            self.broadcast_state()

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

    # ----------------------------------------------------------------------- guards
    def broadcast_state(self, src=0):
        """Parameters, BatchNorm buffers and step counters of rank `src` to every rank."""
        for t in (self.flat_param, self.flat_buf, self.flat_nbt):
            dist.broadcast(t, src)
        self.model._generation += 1

    def check_bound(self):
        """The engine (and its captured CUDA graphs) hold raw pointers into the flat buffers: a later `model.to()`, `.float()`,
        `load_state_dict(assign=True)` or anything else that re-seats `p.data` would silently detach the module from the engine."""
        for k, t, ptr in self._bound:
            if t.data_ptr() != ptr:
                raise RuntimeError(f"pidnet_b200: the storage of '{k}' was replaced after the trainer was created (model.to(), .float(), "
                                   "assign=True ...); the engine still trains the old buffers -- create the model on its device first, "
                                   "or drop model._engine_trainer to rebind")

    def poll_errors(self, wait=False):
        """Deferred report of the last criterion call, without a host sync on the hot path: labels that are neither ignore_label
        nor a class index (the reference's gather faults on the device) and empty OHEM sets (IndexError, utils/criterion.py:73)."""
        if self._pending is None:
            return
        ev, crit = self._pending
        if wait:
            ev.synchronize()
        elif not ev.query():
            return
        self._pending = None
        o = self._host16
        if o[12] > 0:
            raise RuntimeError(f'pidnet_b200: {int(o[12])} label(s) are neither ignore_label nor in [0, num_classes) -- '
                               'set OhemCrossEntropy(ignore_label=...) to the value the dataset uses (Cityscapes: 255); '
                               'those pixels were dropped and the loss poisoned with NaN')
        if crit and (o[8] == 0 or o[9] == 0):
            raise IndexError('index -1 is out of bounds for dimension 0 with size 0')   # what the reference raises

    def _record_out(self, crit):
        self._host16.copy_(self.out16, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._pending = (ev, crit)

    def _plan(self, N, H, W):
        if self.planned != (N, H, W):
            _lib.check(self.lib.pidnet_train_plan(self.h, N, H, W, None))
            self.planned = (N, H, W)
            self._seg_ranges = None

    # ----------------------------------------------------------------------- forward
    def step(self, x, labels, bd_gt, class_weights, crit_cfg, backward=True, want_logits=True, aux_ce_map=None):
        """Train-mode forward + criterion.  backward: True / 1 = also the whole network backward into self.flat_grad (no
        all-reduce); 2 = criterion values and logit gradients only (`backward()` runs the network backward later);
        False / 0 = values only.  Returns (out16, [x_extra_p, x_, x_extra_d])."""
        self.poll_errors()
        self.check_bound()
        x = x.contiguous().float()
        labels = labels.contiguous().long()
        bd_gt = bd_gt.contiguous().float()
        N, _, H, W = x.shape
        with torch.cuda.device(self.device):
            self._plan(N, H, W)
            ncls = self.model._cfg['num_classes']
            outs = [None, None, None]
            if want_logits:
                outs = [torch.empty(N, ncls, H // 8, W // 8, device=self.device), torch.empty(N, ncls, H // 8, W // 8, device=self.device),
                        torch.empty(N, 1, H // 8, W // 8, device=self.device)]
            p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
            stream = torch.cuda.current_stream(self.device).cuda_stream
            cw = self._device_class_weights(class_weights)
            _lib.check(self.lib.pidnet_train_step(self.h, C.c_void_p(stream), p(x), p(labels), p(bd_gt), p(cw), C.byref(crit_cfg),
                                                  int(backward), p(self.out16), p(outs[1]), p(outs[0]), p(outs[2]), p(aux_ce_map)))
            self._record_out(True)
        self._after_forward()
        return self.out16, outs

    def _device_class_weights(self, w):
        """fp32 device copy of the criterion's class weights, cached: the reference keeps them in a CPU tensor
        (datasets/cityscapes.py:55-59), and a `.to(device)` from pageable memory every step is a synchronous copy behind all queued
        work -- a full host/device synchronisation per step that keeps the host from running ahead of the GPU."""
        if w is None:
            return None
        if w.is_cuda and w.dtype == torch.float32 and w.is_contiguous() and w.device == self.device:
            return w
        key = (w.data_ptr(), w._version, tuple(w.shape), w.dtype)
        c = getattr(self, '_cw_cache', None)
        if c is None or c[0] != key or c[1] is not w:
            self._cw_cache = c = (key, w, w.detach().to(self.device, torch.float32).contiguous())
        return c[2]

    def forward_train(self, x):
        """Train-mode forward only (batch statistics, running-stat update): [x_extra_p, x_, x_extra_d]."""
        self.check_bound()
        x = x.contiguous().float()
        N, _, H, W = x.shape
        with torch.cuda.device(self.device):
            self._plan(N, H, W)
            ncls = self.model._cfg['num_classes']
            outs = [torch.empty(N, ncls, H // 8, W // 8, device=self.device), torch.empty(N, ncls, H // 8, W // 8, device=self.device),
                    torch.empty(N, 1, H // 8, W // 8, device=self.device)]
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(self.lib.pidnet_train_forward(self.h, C.c_void_p(stream), C.c_void_p(x.data_ptr()),
                                                     C.c_void_p(outs[1].data_ptr()), C.c_void_p(outs[0].data_ptr()),
                                                     C.c_void_p(outs[2].data_ptr())))
        self._after_forward()
        return outs

    def _after_forward(self):
        self.flat_nbt += 1                                      # nn.BatchNorm2d.num_batches_tracked of every layer
        self.seq += 1
        self.model._generation += 1                             # running statistics changed: eval plans must re-read them

    # ----------------------------------------------------------------------- backward
    def segment_ranges(self):
        """Per backward range k: the [begin, end) float ranges of the flat gradient that are final after it."""
        if self._seg_ranges is None:
            nseg = int(self.lib.pidnet_train_num_segments(self.h))
            out = []
            for s in range(nseg):
                n = C.c_int()
                _lib.check(self.lib.pidnet_train_segment_ranges(self.h, s, C.c_void_p(self.flat_grad.data_ptr()), None, 0, C.byref(n)))
                buf = (C.c_int64 * (2 * max(n.value, 1)))()
                _lib.check(self.lib.pidnet_train_segment_ranges(self.h, s, C.c_void_p(self.flat_grad.data_ptr()), buf, n.value, C.byref(n)))
                out.append([(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n.value)])
            self._seg_ranges = out
        return self._seg_ranges

    def backward(self, x, logit_grads=None, allreduce=True):
        """Network backward of the latest train-mode forward into self.flat_grad (overwritten), from the criterion's own logit
        gradients (`step(..., backward=2)`) or from `logit_grads = (g_x_extra_p, g_x_, g_x_extra_d)`.  With several ranks the
        gradient is averaged (reference: DataParallel's reduce-add + `losses.mean()` over replicas, tools/train.py:136 /
        utils/function.py:44) by all-reduces of the ranges each backward segment has finalised, overlapped with the next one."""
        x = x.contiguous().float()
        gp = gm = gd = None
        if logit_grads is not None:
            gp, gm, gd = [g.contiguous().float() for g in logit_grads]
        p = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        world = _world() if allreduce else 1
        with torch.cuda.device(self.device):
            stream = torch.cuda.current_stream(self.device).cuda_stream
            if world == 1 or not self.overlap_allreduce:
                _lib.check(self.lib.pidnet_train_backward(self.h, C.c_void_p(stream), p(x), p(gm), p(gp), p(gd), -1))
                if world > 1:
                    dist.all_reduce(self.flat_grad[:self.n_param], op=dist.ReduceOp.SUM)
            else:
                from .parallel import BucketedAllReduce
                exchange = BucketedAllReduce(self.flat_grad)
                for seg, ranges in enumerate(self.segment_ranges()):
                    _lib.check(self.lib.pidnet_train_backward(self.h, C.c_void_p(stream), p(x), p(gm), p(gp), p(gd), seg))
                    exchange.launch(ranges)
                exchange.finish()
        return world

    def allreduce_gradients(self, average=True):
        """Sum (mean) the whole flat gradient over all ranks with one all-reduce (for callers that ran `step(backward=True)`)."""
        from .parallel import allreduce_flat_gradient
        allreduce_flat_gradient(self.flat_grad, self.n_param, average)

    def backward_and_publish(self, x, scale, logit_grads=None):
        """`backward()` + hand the parameter gradients to autograd's `.grad` slots.  Returns None (published zero-copy: every
        `p.grad` IS its slice of the flat gradient buffer) or a list of gradient tensors for autograd to accumulate."""
        params, views = self._params, self._views
        alias = [p.grad is not None and p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, views)]
        foreign = any(p.grad is not None and not a for p, a in zip(params, alias))
        # a `.grad` that already is our view holds either zeros (`zero_grad(set_to_none=False)` after the forward, the reference's
        # order utils/function.py:46-47) or an earlier gradient to accumulate onto: keep it, the backward overwrites the buffer
        keep = self.flat_grad.clone() if any(alias) else None
        world = self.backward(x, logit_grads)
        if scale is not None or world > 1:
            s = (scale if scale is not None else 1.0) / world
            self.flat_grad.mul_(s)
        if not foreign:
            if keep is not None:
                self.flat_grad.add_(keep)
            for p, v in zip(params, views):
                p.grad = v
            return None
        new = self.flat_grad.clone()                # some parameter carries a gradient tensor of its own: let autograd accumulate
        if keep is not None:
            self.flat_grad.copy_(keep)
        off, grads = 0, []
        for p in params:
            n = p.numel()
            grads.append(new[off:off + n].view(p.shape))
            off += (n + 3) // 4 * 4
        return grads

    def set_option(self, name, value):
        """Engine option of the training path: "use_graph" (1 = replay CUDA graphs after the first step), "bwd_segments" ..."""
        _lib.check(self.lib.pidnet_train_set_option(self.h, name.encode(), int(value)))
        self._seg_ranges = None

    def debug_tensor(self, name, grad=False):
        shape = (C.c_int64 * 4)()
        _lib.check(self.lib.pidnet_train_debug_tensor(self.h, name.encode(), int(grad), None, shape))
        t = torch.empty(tuple(shape), dtype=torch.float32)
        _lib.check(self.lib.pidnet_train_debug_tensor(self.h, name.encode(), int(grad), C.c_void_p(t.data_ptr()), shape))
        return t


def _check_turn(ctx, what):
    tr = ctx.trainer
    if ctx.done:
        raise RuntimeError(f'pidnet_b200: {what} was already back-propagated (the engine does not retain the graph twice)')
    if ctx.seq != tr.seq:
        raise RuntimeError(f'pidnet_b200: {what} belongs to an earlier train-mode forward; the engine keeps the activations of the '
                           'latest forward only -- call backward() before the next forward of this model')
    ctx.done = True


class _TrainStepFn(torch.autograd.Function):
    """FullModel's train-mode forward: network + fused criterion now, network backward inside `loss.backward()`."""

    @staticmethod
    def forward(ctx, trainer, x, labels, bd_gt, class_weights, crit_cfg, aux_ce_map, *params):
        out16, outs = trainer.step(x, labels, bd_gt, class_weights, crit_cfg, backward=2, aux_ce_map=aux_ce_map)
        ctx.trainer, ctx.x, ctx.seq, ctx.done, ctx.nparam = trainer, x, trainer.seq, False, len(params)
        ctx.mark_non_differentiable(*[o for o in outs])
        return (out16[0:1].clone(), out16.clone(), *outs)

    @staticmethod
    def backward(ctx, g_loss, *unused):
        _check_turn(ctx, 'this loss')
        grads = ctx.trainer.backward_and_publish(ctx.x, g_loss.reshape(()))
        ctx.x = None
        if grads is None:
            return (None,) * (7 + ctx.nparam)
        return (None, None, None, None, None, None, None, *grads)


class _TrainForwardFn(torch.autograd.Function):
    """`PIDNet.forward` in train mode under autograd: three outputs now, the engine backward from the caller's logit gradients
    when a loss built on them is back-propagated (what `self.model(inputs)` is to the reference's FullModel, utils/utils.py:39).
    The input image receives no gradient."""

    @staticmethod
    def forward(ctx, trainer, x, *params):
        outs = trainer.forward_train(x)
        ctx.trainer, ctx.x, ctx.seq, ctx.done, ctx.nparam = trainer, x, trainer.seq, False, len(params)
        return tuple(outs)

    @staticmethod
    def backward(ctx, g_p, g_m, g_d):
        _check_turn(ctx, 'this forward')
        tr = ctx.trainer
        N, _, H, W = ctx.x.shape
        ncls = tr.model._cfg['num_classes']
        z = lambda g, c: g if g is not None else torch.zeros(N, c, H // 8, W // 8, device=tr.device)
        grads = tr.backward_and_publish(ctx.x, None, (z(g_p, ncls), z(g_m, ncls), z(g_d, 1)))
        ctx.x = None
        if grads is None:
            return (None,) * (2 + ctx.nparam)
        return (None, None, *grads)
