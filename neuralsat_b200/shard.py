"""Multi-GPU plumbing of the path: sub-domains are independent units, so the
batch is cut on dim 0 into contiguous slices, one per rank (the split the reference's unused
`BoundDataParallel` does, AL/bound_multi_gpu.py:67,86); weights and the plan are replicated; every
rank bounds its slice with its own alpha/beta/Adam state.  The data path has NO collective.  The two
exchange steps around it use torch.distributed (NCCL over NVLink on the GPUs, gloo in the CPU tests):

  gather_lower_bounds  all_gather of lb[Bd_local,S] so that every rank can take the global prune /
                       stop decision (per DOMAIN, never per rank, so results equal the 1-GPU run)
  rebalance            all_to_all of packed domain records when slice sizes diverge after pruning
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


def slice_bounds(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced slices of range(n): the first n % world ranks get one extra unit."""
    base, extra = divmod(n, world)
    out, start = [], 0
    for r in range(world):
        size = base + (1 if r < extra else 0)
        out.append((start, start + size))
        start += size
    return out


# Fields of a batch / of the reference's AbstractResults whose tensors carry the domain batch on dim 2: the slopes
# [2,S1,Bd,...] (NS/abstractor/utils.py:63-74).  Everything else is batch-leading.
BATCH_DIM2_FIELDS = ('alpha', 'slopes')
# per-domain Python lists (one entry per domain)
PER_DOMAIN_LISTS = ('betas', 'histories', 'sat_solvers')


def shard_tree(obj, lo: int, hi: int, n: int, batch_dim: int = 0, field: str = ''):
    """Slice the domains [lo:hi] out of a batch dict / AbstractResults-like tree.  The batch axis is decided by the
    FIELD NAME, never guessed from shapes: tensors under a key in BATCH_DIM2_FIELDS are cut on dim 2, lists under a key
    in PER_DOMAIN_LISTS are cut as lists, every other tensor on dim 0.  A tensor whose batch axis does not have length
    `n` is an error (a silently mis-sliced input would give wrong bounds, not a crash)."""
    if isinstance(obj, torch.Tensor):
        if obj.dim() <= batch_dim or obj.shape[batch_dim] != n:
            raise ValueError(f"field '{field}': shape {tuple(obj.shape)} has no batch of {n} on dim {batch_dim}")
        return obj.narrow(batch_dim, lo, hi - lo).contiguous()
    if isinstance(obj, dict):
        out = {}
        for k, v in obj.items():
            if k in PER_DOMAIN_LISTS and isinstance(v, (list, tuple)):
                if len(v) != n:
                    raise ValueError(f"field '{k}': {len(v)} entries for {n} domains")
                out[k] = type(v)(v[lo:hi])
            else:
                out[k] = shard_tree(v, lo, hi, n, 2 if k in BATCH_DIM2_FIELDS else batch_dim, k if isinstance(k, str) else field)
        return out
    if isinstance(obj, (list, tuple)):
        return type(obj)(shard_tree(v, lo, hi, n, batch_dim, field) for v in obj)
    return obj


def gather_lower_bounds(lb_local: torch.Tensor, sizes: Sequence[int], dist=None) -> torch.Tensor:
    """lb_local [Bd_r,S] on every rank -> lb [sum Bd_r, S] on every rank, in rank order.
    `sizes[r]` = Bd_r (known to every rank from slice_bounds / the last rebalance)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return lb_local
    world = dist.get_world_size()
    S = lb_local.shape[1]
    if len(set(sizes)) == 1:
        out = torch.empty(world * sizes[0], S, dtype=lb_local.dtype, device=lb_local.device)
        dist.all_gather_into_tensor(out, lb_local.contiguous())
        return out
    mx = max(sizes)
    pad = torch.zeros(mx, S, dtype=lb_local.dtype, device=lb_local.device)
    pad[:lb_local.shape[0]] = lb_local
    out = torch.empty(world * mx, S, dtype=lb_local.dtype, device=lb_local.device)
    dist.all_gather_into_tensor(out, pad)
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)], dim=0)


def transfer_plan(counts: Sequence[int]) -> List[List[int]]:
    """plan[src][dst] = number of records rank src sends to rank dst so that every rank ends with
    floor/ceil(total/world) records; deterministic (every rank computes the same plan), moves the
    minimum number of records, and takes them from the END of the donor's queue."""
    world = len(counts)
    target = [hi - lo for lo, hi in slice_bounds(sum(counts), world)]
    surplus = [c - t for c, t in zip(counts, target)]
    plan = [[0] * world for _ in range(world)]
    donors = [r for r in range(world) if surplus[r] > 0]
    takers = [r for r in range(world) if surplus[r] < 0]
    di = ti = 0
    while di < len(donors) and ti < len(takers):
        d, t = donors[di], takers[ti]
        k = min(surplus[d], -surplus[t])
        plan[d][t] += k
        surplus[d] -= k
        surplus[t] += k
        if surplus[d] == 0:
            di += 1
        if surplus[t] == 0:
            ti += 1
    return plan


def rebalance(records: Dict[str, torch.Tensor], counts: Sequence[int], dist=None) -> Dict[str, torch.Tensor]:
    """records: {field: tensor with the local queue on dim 0}; counts[r] = queue length of rank r
    (gathered beforehand, e.g. with gather_lower_bounds on a [1,1] tensor).  Returns the local queue
    after the exchange: kept records first (original order), then received ones in source-rank order."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return records
    world, rank = dist.get_world_size(), dist.get_rank()
    plan = transfer_plan(counts)
    send = plan[rank]
    recv = [plan[src][rank] for src in range(world)]
    n_local = counts[rank]
    keep = n_local - sum(send)
    out = {}
    for name, t in records.items():
        assert t.shape[0] == n_local, (name, t.shape, n_local)
        row = 1
        for d in t.shape[1:]:
            row *= int(d)
        flat = t.reshape(n_local, row)                 # explicit sizes: a rank whose queue is empty has 0 rows
        send_buf = flat[keep:].contiguous()
        recv_buf = torch.empty(sum(recv), row, dtype=t.dtype, device=t.device)
        dist.all_to_all_single(recv_buf, send_buf, output_split_sizes=recv, input_split_sizes=send)
        out[name] = torch.cat([flat[:keep], recv_buf], dim=0).reshape(keep + sum(recv), *t.shape[1:])
    return out
