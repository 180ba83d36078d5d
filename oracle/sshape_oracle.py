"""CPU oracle for the S-shaped activations (sigmoid / tanh): restatement of
auto_LiRPA/operators/tanh.py (paths relative to /root/reference/neuralsat-pt201, AL = auto_LiRPA).

TEST INFRASTRUCTURE ONLY (see oracle/crown_oracle.py).  Parity status: PINNED by
tests/test_oracle_golden.py against tests/golden/fc_sigmoid.pt and fc_tanh.pt, which were recorded
from the unmodified reference by oracle/gen_golden.py.

  tangent tables d_lower / d_upper ... AL/operators/tanh.py:65-130   (`tables`)
  relaxation lines .................... AL/operators/tanh.py:135-290  (`relax`)
  linear-relaxation bookkeeping ....... AL/operators/activation_base.py:31-60
  backward through the relaxation ..... AL/operators/activation_base.py:247-304 (`backward`)
  sign-split multiply + its backward .. AL/operators/clampmult.py:17-95

Only the lower-bound side (index 0 of the reference's leading "2" dimension) is restated: the BaB
loop never asks for upper bounds.
"""
from __future__ import annotations

import torch

STEP = 0.01
X_LIMIT = 500


def dtanh(x):
    """AL/operators/tanh.py:8-13."""
    mask = torch.lt(torch.abs(x), 25.0).to(x.dtype)
    cosh = torch.cosh(mask * x + 1 - mask)
    return mask * (1. / cosh.pow(2))


def dsigmoid(x):
    """AL/operators/tanh.py:15-16."""
    return torch.sigmoid(x) * (1 - torch.sigmoid(x))


FUNCS = {'tanh': (torch.tanh, dtanh), 'sigmoid': (torch.sigmoid, dsigmoid)}
_TABLES = {}


@torch.no_grad()
def tables(op: str):
    """AL/operators/tanh.py:65-130 -> (d_lower, d_upper), fp32 [50005] each."""
    if op in _TABLES:
        return _TABLES[op]
    func, dfunc = FUNCS[op]
    num = int(X_LIMIT / STEP)
    max_iter = 100

    def check_lower(upper, d):
        k = dfunc(d)
        return k * (upper - d) + func(d) <= func(upper)

    def check_upper(lower, d):
        k = dfunc(d)
        return k * (lower - d) + func(d) >= func(lower)

    upper = STEP * torch.arange(0, num + 5)
    r = torch.zeros_like(upper)
    l = -torch.ones_like(upper)
    while True:
        checked = check_lower(upper, l).int()
        l = checked * l + (1 - checked) * (l * 2)
        if checked.sum() == l.numel():
            break
    for _ in range(max_iter):
        m = (l + r) / 2
        checked = check_lower(upper, m).int()
        l = checked * m + (1 - checked) * l
        r = checked * r + (1 - checked) * m
    d_lower = l.clone()

    lower = -STEP * torch.arange(0, num + 5)
    l = torch.zeros_like(upper)
    r = torch.ones_like(upper)
    while True:
        checked = check_upper(lower, r).int()
        r = checked * r + (1 - checked) * (r * 2)
        if checked.sum() == l.numel():
            break
    for _ in range(max_iter):
        m = (l + r) / 2
        checked = check_upper(lower, m).int()
        l = (1 - checked) * m + checked * l
        r = (1 - checked) * r + checked * m
    d_upper = r.clone()
    _TABLES[op] = (d_lower, d_upper)
    return _TABLES[op]


def lookup(op, lower, upper):
    """AL/operators/tanh.py:150-187: tangent points that are valid for [lower, upper]."""
    d_lower_t, d_upper_t = tables(op)
    index = torch.max(torch.zeros(upper.numel(), dtype=torch.long),
                      (upper / STEP).to(torch.long).reshape(-1)) + 1
    d_lower = torch.where((index < d_lower_t.numel()).view(lower.shape),
                          torch.index_select(d_lower_t, 0, index.clamp(max=d_lower_t.numel() - 1)).view(lower.shape),
                          lower)
    index = torch.max(torch.zeros(lower.numel(), dtype=torch.long),
                      (lower / -STEP).to(torch.long).reshape(-1)) + 1
    d_upper = torch.where((index < d_upper_t.numel()).view(upper.shape),
                          torch.index_select(d_upper_t, 0, index.clamp(max=d_upper_t.numel() - 1)).view(upper.shape),
                          upper)
    return d_lower, d_upper


def init_alpha(op, lower, upper, S1=1):
    """AL/operators/tanh.py:54-63 (`_init_opt_parameters_impl`): [8,S1,Bd,n]."""
    d_lower, d_upper = lookup(op, lower, upper)
    alpha = torch.empty(8, S1, *lower.shape)
    alpha[:4] = (lower + upper) / 2
    alpha[4:6] = d_lower
    alpha[6:8] = d_upper
    return alpha


def relax(op, lower, upper, alpha):
    """AL/operators/tanh.py:135-290 -> lw, lb, uw, ub.

    alpha: None (plain CROWN: middle-point tangents / table tangents) or the parameter tensor
    [8,S1,Bd,n]; its `.data` is clipped IN PLACE like the reference does (:191-198).
    Shapes: [Bd,n] without alpha, [S1,Bd,n] with alpha (index 0 of the reference's leading 2)."""
    func, dfunc = FUNCS[op]
    mask_pos = lower >= 0
    mask_neg = upper <= 0
    mask_both = torch.logical_not(torch.logical_or(mask_pos, mask_neg))
    y_l, y_u = func(lower), func(upper)
    k_direct = (y_u - y_l) / (upper - lower).clamp(min=1e-8)
    d_lower, d_upper = lookup(op, lower, upper)

    if alpha is not None:
        S1 = alpha.shape[1]
        shape = (S1, *lower.shape)
    else:
        shape = lower.shape
    lw = torch.zeros(shape)
    lb = torch.zeros(shape)
    uw = torch.zeros(shape)
    ub = torch.zeros(shape)

    def add(w_out, b_out, mask, k, x0, y0=None):
        """activation_base.py:31-60 with a mask; returns the updated (w, b)."""
        if y0 is None:
            y0 = func(x0)
        b = -x0 * k + y0
        return torch.where(mask, k.expand(shape), w_out), torch.where(mask, b.expand(shape), b_out)

    uw, ub = add(uw, ub, mask_neg, k_direct, lower, y_l)
    lw, lb = add(lw, lb, mask_pos, k_direct, lower, y_l)
    if alpha is not None:
        with torch.no_grad():
            alpha.data[0:2] = torch.max(torch.min(alpha[0:2], upper), lower)
            alpha.data[2:4] = torch.max(torch.min(alpha[2:4], upper), lower)
            alpha.data[4:6] = torch.min(alpha[4:6], d_lower)
            alpha.data[6:8] = torch.max(alpha[6:8], d_upper)
        tp_pos, tp_neg, tp_both_lower, tp_both_upper = alpha[0], alpha[2], alpha[4], alpha[6]
        mask_direct = torch.logical_and(mask_both, k_direct < dfunc(lower))
        lw, lb = add(lw, lb, mask_direct, k_direct, lower, y_l)
        lw, lb = add(lw, lb, torch.logical_xor(mask_both, mask_direct), dfunc(tp_both_lower), tp_both_lower)
        mask_direct = torch.logical_and(mask_both, k_direct < dfunc(upper))
        uw, ub = add(uw, ub, mask_direct, k_direct, lower, y_l)
        uw, ub = add(uw, ub, torch.logical_xor(mask_both, mask_direct), dfunc(tp_both_upper), tp_both_upper)
        lw, lb = add(lw, lb, mask_neg, dfunc(tp_neg), tp_neg)
        uw, ub = add(uw, ub, mask_pos, dfunc(tp_pos), tp_pos)
    else:
        m = (lower + upper) / 2
        y_m = func(m)
        k = dfunc(m)
        lw, lb = add(lw, lb, mask_neg, k, m, y_m)
        uw, ub = add(uw, ub, mask_pos, k, m, y_m)
        k = dfunc(d_lower)
        mask_direct = torch.logical_and(mask_both, k_direct < dfunc(lower))
        lw, lb = add(lw, lb, mask_direct, k_direct, lower, y_l)
        lw, lb = add(lw, lb, torch.logical_xor(mask_both, mask_direct), k, d_lower)
        k = dfunc(d_upper)
        mask_direct = torch.logical_and(mask_both, k_direct < dfunc(upper))
        uw, ub = add(uw, ub, mask_direct, k_direct, lower, y_l)
        uw, ub = add(uw, ub, torch.logical_xor(mask_both, mask_direct), k, d_upper)
    return lw, lb, uw, ub


class _SignSplit4(torch.autograd.Function):
    """AL/operators/clampmult.py:17-95 with both bias terms (the `A >= 0` tie rule of its backward)."""

    @staticmethod
    def forward(ctx, A, d_pos, d_neg, b_pos, b_neg):
        ctx.save_for_backward(A, d_pos, d_neg, b_pos, b_neg)
        A_pos = A.clamp(min=0)
        A_neg = A.clamp(max=0)
        A_new = d_pos * A_pos + d_neg * A_neg
        bias = torch.einsum('sb...,sb...->sb', A_pos, b_pos.expand_as(A_pos)) + \
            torch.einsum('sb...,sb...->sb', A_neg, b_neg.expand_as(A_neg))
        return A_new, bias

    @staticmethod
    def backward(ctx, gA_out, gbias):
        A, d_pos, d_neg, b_pos, b_neg = ctx.saved_tensors
        gbias = gbias.view(gbias.shape + (1,) * (A.dim() - gbias.dim()))
        pos = (A >= 0).to(gA_out.dtype)
        neg = 1. - pos
        pg, ng = pos * gA_out, neg * gA_out
        pb, nb = pos * gbias, neg * gbias
        gA = d_pos * pg + d_neg * ng + b_pos * pb + b_neg * nb

        def _reduce(g, ref):
            while g.dim() > ref.dim():
                g = g.sum(0)
            for i, (a, b) in enumerate(zip(g.shape, ref.shape)):
                if a != b:
                    g = g.sum(i, keepdim=True)
            return g
        return gA, _reduce(A * pg, d_pos), _reduce(A * ng, d_neg), _reduce(A * pb, b_pos), _reduce(A * nb, b_neg)


def backward(op, A, lower, upper, alpha):
    """AL/operators/activation_base.py:247-304 (lower side): A [S,Bd,*shape] -> (A_pre, bias [S,Bd])."""
    lw, lb, uw, ub = relax(op, lower, upper, alpha)
    if alpha is None:
        lw, lb, uw, ub = (t.unsqueeze(0) for t in (lw, lb, uw, ub))
    return _SignSplit4.apply(A, lw.expand_as(A).contiguous(), uw.expand_as(A).contiguous(),
                             lb.expand_as(A).contiguous(), ub.expand_as(A).contiguous())
