"""CPU tests of the data-parallel host logic (world_size 2, gloo): the gradient all-reduce of the trainer and the row
sharding of inference.  No GPU."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from vla_touch_b200.optim import allreduce_gradients, bucket_plan


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    grads = [torch.randn(s, generator=g) for s in ((7, 5), (33,), (4, 4, 3), (1,))]
    g0, g1 = torch.Generator().manual_seed(100), torch.Generator().manual_seed(101)
    ref = [torch.randn(s, generator=g0) + torch.randn(s, generator=g1) for s in ((7, 5), (33,), (4, 4, 3), (1,))]
    w = allreduce_gradients(grads, bucket_elems=40)       # forces several buckets
    ok = w == world and all(torch.allclose(a, b, atol=1e-6) for a, b in zip(grads, ref))
    # inference shards rows with no collective: each rank owns rows [rank*B/world, (rank+1)*B/world)
    rows = torch.arange(10).chunk(world)[rank]
    gathered = [torch.zeros_like(rows) for _ in range(world)]
    dist.all_gather(gathered, rows)
    ok = ok and torch.equal(torch.cat(gathered), torch.arange(10))
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_bucket_plan_covers_every_tensor_once():
    numels = [5, 100, 7, 64, 1, 300]
    b = bucket_plan(numels, bucket_elems=128)
    flat = [i for x in b for i in x]
    assert sorted(flat) == list(range(len(numels)))
    assert flat == list(reversed(range(len(numels))))          # reverse registration order
    assert all(sum(numels[i] for i in x) <= 128 or len(x) == 1 for x in b)


def test_gradient_allreduce_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}


def test_allreduce_is_identity_without_process_group():
    g = [torch.ones(3)]
    assert allreduce_gradients(g) == 1 and torch.equal(g[0], torch.ones(3))
