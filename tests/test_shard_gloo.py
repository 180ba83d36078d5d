"""CPU, world_size 2 over gloo: the N>1 plumbing of the path (neuralsat_b200/shard.py).  Sharding the
domain batch, bounding each slice independently and gathering lb must equal the single-rank result;
the work-queue rebalance must preserve the multiset of records and even out the queue lengths."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralsat_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_bound(batch):
    """stands in for the per-slice bounding: any row-wise function of the domain's own data"""
    return (batch['C'].sum(-1) * batch['x_L'].mean(-1, keepdim=True) - batch['alpha'][0, 0].sum(-1, keepdim=True))


def _make(n):
    g = torch.Generator().manual_seed(3)
    return {'C': torch.randn(n, 1, 5, generator=g), 'x_L': torch.rand(n, 7, generator=g),
            'alpha': torch.rand(2, 1, n, 4, generator=g), 'betas': [{'k': i} for i in range(n)],
            'lower': [torch.randn(n, 6, generator=g), torch.randn(n, 3, generator=g)]}


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        full = _make(n)
        bounds = shard.slice_bounds(n, world)
        lo, hi = bounds[rank]
        mine = shard.shard_tree(full, lo, hi, n)
        assert mine['alpha'].shape == (2, 1, hi - lo, 4) and len(mine['betas']) == hi - lo
        assert mine['lower'][1].shape == (hi - lo, 3)
        lb = _fake_bound(mine)
        sizes = [b - a for a, b in bounds]
        glb = shard.gather_lower_bounds(lb, sizes, dist)
        ok_gather = torch.equal(glb, _fake_bound(full))
        # prune with a global, per-domain rule, then rebalance the survivors
        keep = glb[lo:hi, 0] > glb[:, 0].median()
        rec = {'id': torch.arange(lo, hi)[keep], 'x_L': mine['x_L'][keep], 'lb': lb[keep]}
        cnt = torch.tensor([int(keep.sum())])
        allc = [torch.zeros(1, dtype=torch.long) for _ in range(world)]
        dist.all_gather(allc, cnt)
        counts = [int(c) for c in allc]
        new = shard.rebalance(rec, counts, dist)
        q.put((rank, ok_gather, counts, new['id'].tolist(), new['x_L'].shape, new['lb'].flatten().tolist()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [10, 7])
def test_shard_gather_rebalance_world2(n):
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = sorted([q.get() for _ in range(world)])
    full = _make(n)
    glb = _fake_bound(full)
    survivors = (glb[:, 0] > glb[:, 0].median()).nonzero().flatten().tolist()
    ids = []
    for rank, ok, counts, new_ids, xshape, lbs in res:
        assert ok
        assert xshape[0] == len(new_ids) and xshape[1] == 7
        for i, v in zip(new_ids, lbs):
            assert abs(v - float(glb[i, 0])) < 1e-6            # records stay intact
        ids += new_ids
    assert sorted(ids) == survivors                             # nothing lost, nothing duplicated
    lens = [len(r[3]) for r in res]
    assert max(lens) - min(lens) <= 1                           # evened out


def _worker_empty(rank, world, port, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        # rank 1 pruned every domain: the case rebalance exists for
        n_local = 6 if rank == 0 else 0
        rec = {'id': torch.arange(n_local), 'lower': torch.arange(n_local * 12, dtype=torch.float32).reshape(n_local, 3, 4),
               'lb': torch.arange(n_local, dtype=torch.float32).reshape(n_local, 1)}
        new = shard.rebalance(rec, [6, 0], dist)
        q.put((rank, new['id'].tolist(), tuple(new['lower'].shape), new['lower'].flatten().tolist()))
    finally:
        dist.destroy_process_group()


def test_rebalance_with_an_empty_rank():
    world, port = 2, _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker_empty, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    res = dict((r[0], r[1:]) for r in [q.get() for _ in range(world)])
    assert res[0][0] == [0, 1, 2] and res[1][0] == [3, 4, 5]
    assert res[0][1] == (3, 3, 4) and res[1][1] == (3, 3, 4)
    assert res[1][2] == [float(v) for v in range(36, 72)]


def test_shard_tree_uses_field_names_not_shapes():
    """n == 2 domains: batch-leading tensors whose other dims happen to be 2 must still be cut on dim 0."""
    n = 2
    batch = {'C': torch.arange(4.).reshape(2, 1, 2), 'lower': [torch.arange(16.).reshape(2, 4, 2, 1)],
             'alpha': [torch.arange(12.).reshape(2, 1, 2, 3)], 'betas': [{'a': 0}, {'a': 1}]}
    part = shard.shard_tree(batch, 1, 2, n)
    assert part['C'].shape == (1, 1, 2) and torch.equal(part['C'], batch['C'][1:2])
    assert part['lower'][0].shape == (1, 4, 2, 1)
    assert part['alpha'][0].shape == (2, 1, 1, 3) and torch.equal(part['alpha'][0], batch['alpha'][0][:, :, 1:2])
    assert part['betas'] == [{'a': 1}]
    with pytest.raises(ValueError):
        shard.shard_tree({'C': torch.zeros(3, 1, 2)}, 0, 1, n)


def test_transfer_plan_properties():
    for counts in ([5, 0], [0, 9, 1, 2], [3, 3, 3], [100, 1, 1, 1, 1, 1, 1, 1], [0, 0]):
        plan = shard.transfer_plan(counts)
        w = len(counts)
        after = [counts[r] - sum(plan[r]) + sum(plan[s][r] for s in range(w)) for r in range(w)]
        assert sum(after) == sum(counts) and max(after) - min(after) <= 1
        assert all(plan[r][r] == 0 for r in range(w))
        moved = sum(map(sum, plan))
        assert moved == sum(max(0, c - t) for c, t in zip(counts, after))     # minimal movement


def test_slice_bounds():
    assert shard.slice_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert shard.slice_bounds(2, 4) == [(0, 1), (1, 2), (2, 2), (2, 2)]
