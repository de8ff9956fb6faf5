"""Multi-GPU plumbing for the PIDNet path: one process per GPU (torch.distributed), images are
independent units so inference shards the batch with NO data-path collective.
Only bookkeeping (barriers, max-over-ranks timing, optional result gather for tests) uses the
process group.  The reference's single-process nn.DataParallel (tools/train.py:136) is what this
replaces."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def shard_range(num_items, world_size, rank):
    """Contiguous, balanced [begin, end) slice of `num_items` for `rank` (sizes differ by at most 1)."""
    base, extra = divmod(num_items, world_size)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def env_rank():
    return int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('LOCAL_RANK', '0'))


def reduce_max(value, device=None):
    """Max of a python float over all ranks (1 rank: identity)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or 'cpu')
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_flat_gradient(flat, numel=None, average=True):
    """The ONE exchange step of data-parallel training: sum the flat fp32 gradient buffer
    over all ranks and divide by the world size -- what nn.DataParallel's reduce-add to GPU 0 followed by
    `losses.mean()` over replicas computes (tools/train.py:136, utils/function.py:44).  In place; 1 rank: identity.
    Backend-agnostic (NCCL over NVLink on GPUs; gloo in the CPU tests)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return flat
    view = flat if numel is None else flat[:numel]
    dist.all_reduce(view, op=dist.ReduceOp.SUM)
    if average:
        view.div_(dist.get_world_size())
    return flat


class BucketedAllReduce:
    """Gradient exchange overlapped with the backward pass ("bucketed in reverse layer order"): the
    backward runs in consecutive ranges; after range k the caller passes the [begin, end) float ranges of the flat gradient
    that range finalised and `launch` starts one ASYNC all-reduce per range, which proceeds on the communication stream while
    range k + 1 computes.  `finish` makes the caller's stream wait for all of them (the 1/world averaging is folded into
    the caller's final scaling kernel).  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests."""

    def __init__(self, flat):
        self.flat = flat
        self.works = []
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1

    def launch(self, ranges):
        if self.world == 1:
            return
        for b, e in ranges:
            if e > b:
                self.works.append(dist.all_reduce(self.flat[b:e], op=dist.ReduceOp.SUM, async_op=True))

    def finish(self):
        for w in self.works:
            w.wait()
        self.works = []
        return self.world


class ShardedInference:
    """Runs `model` on this rank's slice of a global batch.  `gather=True` (tests / small batches only)
    all-gathers the logits so every rank sees the full result in the original order."""

    def __init__(self, model):
        self.model = model
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1

    def local_slice(self, global_batch):
        b, e = shard_range(global_batch.shape[0], self.world, self.rank)
        return global_batch[b:e]

    def __call__(self, global_batch, gather=False):
        x = self.local_slice(global_batch)
        y = self.model(x) if x.shape[0] > 0 else None
        if not gather or self.world == 1:
            return y
        sizes = [shard_range(global_batch.shape[0], self.world, r) for r in range(self.world)]
        outs = [None] * self.world
        dist.all_gather_object(outs, None if y is None else [t.cpu() for t in (y if isinstance(y, list) else [y])])
        parts = [o for o in outs if o is not None]
        merged = [torch.cat([p[i] for p in parts]) for i in range(len(parts[0]))]
        assert merged[0].shape[0] == sizes[-1][1]
        return merged if isinstance(y, list) or len(merged) > 1 else merged[0]
