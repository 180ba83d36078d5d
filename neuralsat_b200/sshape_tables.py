"""Tangent-point tables of the S-shaped activations, built once per process on the host.

The reference pre-computes, for every pre-activation bound on a 0.01 grid up to abs(x)=500, the tangent
point whose tangent line stays on one side of the function over the whole interval
(auto_LiRPA/operators/tanh.py:65-130: doubling search for a valid start, then 100 bisection steps in
fp32).  The CUDA relaxation kernels (csrc/crown_sshape.cu) index these tables exactly like the
reference indexes its own (`index = max(0, int(bound / 0.01)) + 1`, operators/tanh.py:150-187), so the
tables are produced with the same fp32 torch arithmetic on the CPU and uploaded; they are plan
constants, not part of the per-call path.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

GRID = 0.01
LIMIT = 500
_cache: Dict[Tuple[str, str], Tuple[torch.Tensor, torch.Tensor]] = {}


def _d_tanh(x):
    inside = (x.abs() < 25.0).to(x.dtype)
    return inside * (1. / torch.cosh(inside * x + 1 - inside).pow(2))


def _d_sigmoid(x):
    return torch.sigmoid(x) * (1 - torch.sigmoid(x))


_FN = {'tanh': (torch.tanh, _d_tanh), 'sigmoid': (torch.sigmoid, _d_sigmoid)}


@torch.no_grad()
def _bisect(f, df, anchor, side):
    """side=-1: for anchor >= 0 find the tangent point t <= 0 closest to 0 whose tangent is below f at
    anchor; side=+1: mirror image (anchor <= 0, tangent above f).  Returns the valid end of the bracket."""
    def valid(t):
        line = df(t) * (anchor - t) + f(t)
        return (line <= f(anchor)) if side < 0 else (line >= f(anchor))

    good = torch.full_like(anchor, float(side))       # start at -1 / +1 and double until valid
    while True:
        ok = valid(good).int()
        good = ok * good + (1 - ok) * (good * 2)
        if int(ok.sum()) == good.numel():
            break
    bad = torch.zeros_like(anchor)
    for _ in range(100):
        mid = ((good + bad) / 2) if side < 0 else ((bad + good) / 2)
        ok = valid(mid).int()
        good, bad = ok * mid + (1 - ok) * good, ok * bad + (1 - ok) * mid
    return good


def tangent_tables(op: str, device) -> Tuple[torch.Tensor, torch.Tensor]:
    """-> (d_lower, d_upper) fp32 [50005] on `device` for op in {'sigmoid','tanh'}."""
    key = (op, str(device))
    if key not in _cache:
        f, df = _FN[op]
        n = int(LIMIT / GRID) + 5
        grid = GRID * torch.arange(0, n)
        d_lower = _bisect(f, df, grid, -1)
        d_upper = _bisect(f, df, -GRID * torch.arange(0, n), +1)
        _cache[key] = (d_lower.to(device).contiguous(), d_upper.to(device).contiguous())
    return _cache[key]


def lookup_points(op: str, lower: torch.Tensor, upper: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Table tangent points valid on [lower, upper] (auto_LiRPA/operators/tanh.py:150-187); used to
    initialise the optimisable tangent points (operators/tanh.py:54-63)."""
    d_lower_t, d_upper_t = tangent_tables(op, lower.device)
    n = d_lower_t.numel()
    iu = (upper / GRID).to(torch.long).clamp(min=0) + 1
    il = (lower / -GRID).to(torch.long).clamp(min=0) + 1
    d_lower = torch.where(iu < n, d_lower_t[iu.clamp(max=n - 1)], lower)
    d_upper = torch.where(il < n, d_upper_t[il.clamp(max=n - 1)], upper)
    return d_lower, d_upper
