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


def _train_worker(rank, world, port, out):
    """One data-parallel training step as SURVEY 8e describes it: each rank runs get_loss(...).backward() on its own rows
    (the native program is replaced by the CPU descriptor interpreter), ONE bucketed SUM all-reduce of the gradients, scale by
    1 / world.  Result on every rank == the gradients of the un-sharded batch."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here]
    import bwd_cases
    import plan_emu
    import vt_testutil as U
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    Plan.compile = lambda self: _Interp(self)
    torch.set_num_threads(2)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, T, B = 10, 16, 4
    g = torch.Generator().manual_seed(7)
    x0, x1 = torch.rand(B, T, A, generator=g) * 2 - 1, torch.rand(B, T, A, generator=g) * 2 - 1
    cond, step, z = torch.randn(B, 256, generator=g), torch.rand(B, generator=g), torch.randn(B, T, A, generator=g)

    def grads_of(rows):
        si = bwd_cases.interpolant(A, T, torch.device("cpu"))
        si.step_override, si.z_override = step[rows], z[rows]
        loss, _ = si.get_loss({"obs_cond": cond[rows], "expert_act": x1[rows], "vla_act": x0[rows]}, "cpu")
        loss.backward()
        return [p.grad for p in si.net.parameters()], float(loss.detach())

    per = B // world
    mine, my_loss = grads_of(slice(rank * per, (rank + 1) * per))
    w = allreduce_gradients(mine, bucket_elems=8 << 20)
    for t in mine:
        t.mul_(1.0 / w)
    ok = w == world
    if rank == 0:
        full, full_loss = grads_of(slice(0, B))
        worst = max(float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12) for a, b in zip(mine, full))
        ok = ok and worst <= 3e-2
        out["worst"] = worst
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_data_parallel_training_step_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_train_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        res = dict(out)
        assert res.get(0) is True and res.get(1) is True, res


def test_gradient_arena_buckets_are_complete_when_their_op_index_is_reached():
    """The trainer launches the all-reduce of a bucket of the gradient arena as soon as the backward program has executed the
    bucket's `ready` op count (optim.arena_buckets).  Claim checked here on the CPU descriptor interpreter: after ops [0, ready)
    every element of the bucket already holds its final value (no later op writes into it), the buckets tile the arena, and all
    438 parameter gradients live inside it."""
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path[:0] = [here]
    import plan_emu
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.optim import arena_buckets
    from vla_touch_b200.params import sub_state_dict
    from vla_touch_b200.unet_train import LossBackwardProgram
    A, B, T = 7, 2, 8
    full = syn.synth_state_dict(shp.si_net_shapes(A, 256), 21, prefix="net.")
    lp = LossBackwardProgram([sub_state_dict(full, p) for p in ("b_net.", "v_net.", "s_net.")], A, B, T, 0.03, "cpu")
    g = torch.Generator().manual_seed(3)
    lp.set_inputs(torch.rand(B, T, A, generator=g) * 2 - 1, torch.rand(B, T, A, generator=g) * 2 - 1, torch.randn(B, 256, generator=g),
                  torch.rand(B, generator=g), torch.randn(B, T, A, generator=g))
    arena, allocs = lp.grad_arena()
    buckets = arena_buckets(allocs, arena.numel(), bucket_elems=12 << 20)
    assert len(buckets) >= 4 and buckets[0][0] == 0 and buckets[-1][1] == arena.numel() and buckets[-1][2] is None
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))
    ready = [b[2] for b in buckets[:-1]]
    assert ready == sorted(ready) and all(0 < r < len(lp.plan) for r in ready)
    base, end = arena.data_ptr(), arena.data_ptr() + arena.numel() * 4
    src = lp.grad_sources()
    assert len(src) == 438 and all(base <= t.data_ptr() < end and t.is_contiguous() for t, _, _, _ in src.values())
    snaps, op0 = [], 0
    for a, b, r in buckets:
        r = len(lp.plan) if r is None else r
        plan_emu.run(lp.plan, op0, r - op0)
        op0 = r
        snaps.append(arena[a:b].clone())
    assert float(arena.abs().max()) > 0
    for (a, b, _), snap in zip(buckets, snaps):
        assert torch.equal(arena[a:b], snap)


def _arena_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from vla_touch_b200.optim import allreduce_arena, arena_buckets
    arena = torch.arange(1000, dtype=torch.float32) * (rank + 1)
    allocs = [(i * 100, 100, i) for i in range(10)]
    works = [allreduce_arena(arena[a:b]) for a, b, _ in arena_buckets(allocs, 1000, bucket_elems=250)]
    for w in works:
        w.wait()
    out[rank] = bool(torch.equal(arena, torch.arange(1000, dtype=torch.float32) * 3))
    dist.destroy_process_group()


def test_arena_allreduce_in_place_world2_gloo():
    world = 2
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_arena_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
