"""world_size-2 `gloo` test of the N > 1 path's host logic (no GPU): batch sharding covers the global
batch exactly once with no data-path collective, the max-over-ranks timing reduction works, and the
gathered result equals the single-process result."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pidnet_b200 import parallel as PAR


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _fake_model(x):
    # stands in for PIDNet.forward (which needs a GPU): per-image, batch-independent function
    return [x.mean(dim=(1, 2, 3), keepdim=True) * 2.0, x.amax(dim=(1,), keepdim=True)]


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        assert PAR.env_rank() == (rank, world, rank)
        g = torch.randn(5, 3, 8, 8, generator=torch.Generator().manual_seed(0))   # odd batch: 3 + 2
        runner = PAR.ShardedInference(_fake_model)
        local = runner.local_slice(g)
        assert local.shape[0] == (3 if rank == 0 else 2)
        merged = runner(g, gather=True)
        ref = _fake_model(g)
        ok = all(torch.equal(a, b) for a, b in zip(merged, ref))
        mx = PAR.reduce_max(10.0 + rank)
        # the training exchange step: every rank ends with the MEAN of the per-rank flat gradients; the padding tail of
        # the buffer (beyond numel) is not touched
        flat = torch.arange(10, dtype=torch.float32) * (rank + 1)
        PAR.allreduce_flat_gradient(flat, numel=8, average=True)
        want = torch.arange(10, dtype=torch.float32) * 1.5
        want[8:] = torch.arange(8, 10, dtype=torch.float32) * (rank + 1)
        ok = ok and torch.equal(flat, want)
        summed = PAR.allreduce_flat_gradient(torch.ones(4) * (rank + 1), average=False)
        ok = ok and torch.equal(summed, torch.full((4,), 3.0))
        # bucketed exchange: ranges finalised by successive backward segments are reduced asynchronously, out of order w.r.t.
        # the buffer layout; untouched gaps (other segments' ranges, alignment padding) stay local until their turn
        flat = torch.arange(32, dtype=torch.float32) * (rank + 1)
        ex = PAR.BucketedAllReduce(flat)
        ex.launch([(20, 28), (4, 8)])
        ex.launch([(0, 4), (8, 20)])
        assert ex.finish() == world
        flat[:28] /= world
        want = torch.arange(32, dtype=torch.float32) * 1.5
        want[28:] = torch.arange(28, 32, dtype=torch.float32) * (rank + 1)
        ok = ok and torch.equal(flat, want) and not ex.works
        ret[rank] = (ok, mx)
    finally:
        dist.destroy_process_group()


def test_sharded_inference_gloo_world2():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, mx = ret[r]
        assert ok and mx == 11.0
