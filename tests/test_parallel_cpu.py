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
