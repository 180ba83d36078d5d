"""CPU oracle: plain-PyTorch fp32 restatement of NeuralSAT's theory-solver hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in `neuralsat_b200/` imports this file; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs do, and only
as the checker / the timed CPU baseline.  The product path is the CUDA extension.

Parity status: PINNED.  `tests/test_oracle_golden.py` checks this file against
  * the known-answer vectors of the reference's fixed-weight toy net
    (NS/example/test_model.py:80-108, values in SURVEY.md section 8c), and
  * golden fixtures under tests/golden/ produced by running the UNMODIFIED reference
    (auto_LiRPA BoundedModule.compute_bounds through NetworkAbstractor) in the build
    container with oracle/gen_golden.py.

What is restated (paths relative to /root/reference/neuralsat-pt201, AL = auto_LiRPA):
  backward traversal ............ AL/backward_bound.py:102-311   (`crown_pass`)
  Linear  A.W, A.b .............. AL/operators/linear.py:167-175
  Conv    conv_transpose2d(A) ... AL/operators/convolution.py:51-96
  BatchNorm ..................... AL/operators/normalization.py:103-128
  Add / fan-in accumulation ..... AL/operators/add_sub.py:19-29, AL/backward_bound.py:691-709
  ReLU relaxation ............... AL/operators/relu.py:456-557
  sparse alpha reconstruct ...... AL/operators/relu.py:208-221
  sign-split multiply ........... AL/operators/clampmult.py:17-43 (backward :49-95 via autograd)
  beta injection ................ AL/beta_crown.py:163-204
  S-shaped relaxations .......... AL/operators/tanh.py:135-290 (see `oracle/sshape_oracle.py`)
  concretise .................... AL/perturbations.py:154-183
  alpha/beta optimisation loop .. AL/optimized_bounds.py:255-629 (`optimize`)
  stop criterion ................ AL/utils.py:87-93

The network is given as a list of node dicts in topological order (see `neuralsat_b200.graph`
for the producer; fixtures store the same list):
  {'op': 'input',  'shape': (...)}
  {'op': 'linear', 'in': [i], 'weight': [out,in], 'bias': [out] | None, 'shape': (out,)}
  {'op': 'conv2d', 'in': [i], 'weight', 'bias', 'stride', 'padding', 'dilation', 'groups', 'shape': (C,H,W)}
  {'op': 'batchnorm2d', 'in': [i], 'weight','bias','mean','var','eps', 'shape'}
  {'op': 'add' | 'sub', 'in': [i, j], 'shape'}
  {'op': 'flatten', 'in': [i], 'shape': (n,)}
  {'op': 'addconst', 'in': [i], 'value': [*shape], 'shape'}      (Add/Sub with an unperturbed operand)
  {'op': 'relu', 'in': [i], 'shape'}
The last node is the output node.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------
# elementwise pieces
# ----------------------------------------------------------------------------------------------

def relu_upper_line(lower: torch.Tensor, upper: torch.Tensor):
    """AL/operators/relu.py:456-469 (leaky_alpha = 0)."""
    lb_r = lower.clamp(max=0)
    ub_r = upper.clamp(min=0)
    ub_r = torch.max(ub_r, lb_r + 1e-8)
    upper_d = ub_r / (ub_r - lb_r)
    upper_b = -lb_r * upper_d
    return upper_d, upper_b


def relu_lower_slope(lower, upper, alpha_full: Optional[torch.Tensor], upper_d):
    """alpha given: AL/operators/relu.py:471-494 (_relu_mask_alpha); else adaptive :347-371."""
    if alpha_full is None:
        return (upper_d > 0.5).to(upper_d).unsqueeze(0)
    lower_mask = (lower >= 0).to(lower.dtype)
    upper_mask = (upper <= 0).to(lower.dtype)
    no_mask = (1. - lower_mask) * (1. - upper_mask)
    return torch.clamp(alpha_full, min=0., max=1.) * no_mask + lower_mask


def reconstruct_full_alpha(sparse_alpha: torch.Tensor, n: int, alpha_index: Optional[torch.Tensor]):
    """AL/operators/relu.py:208-221; sparse_alpha [S1,Bd,n_alpha] -> [S1,Bd,n] (flattened neuron dim)."""
    if alpha_index is None:
        return sparse_alpha.reshape(sparse_alpha.shape[0], sparse_alpha.shape[1], n)
    full = torch.zeros(sparse_alpha.shape[0], sparse_alpha.shape[1], n, dtype=sparse_alpha.dtype)
    full[:, :, alpha_index] = sparse_alpha
    return full


def sign_split_multiply(A, d_pos, d_neg, b_pos, b_neg):
    """AL/operators/clampmult.py:17-43 with reduce_bias=True; A [S,Bd,n]."""
    A_pos = A.clamp(min=0)
    A_neg = A.clamp(max=0)
    A_new = d_pos * A_pos + d_neg * A_neg
    bias = 0.
    if b_pos is not None:
        bias = bias + torch.einsum('sbn,sbn->sb', A_pos, b_pos.expand_as(A_pos))
    if b_neg is not None:
        bias = bias + torch.einsum('sbn,sbn->sb', A_neg, b_neg.expand_as(A_neg))
    return A_new, bias


class _SignSplit(torch.autograd.Function):
    """Same forward as `sign_split_multiply` for the ReLU case (b_pos=None), with the
    reference's hand-written backward (AL/operators/clampmult.py:49-95): the A>=0 tie rule."""

    @staticmethod
    def forward(ctx, A, d_pos, d_neg, b_neg):
        ctx.save_for_backward(A, d_pos, d_neg, b_neg)
        A_pos = A.clamp(min=0)
        A_neg = A.clamp(max=0)
        A_new = d_pos * A_pos + d_neg * A_neg
        bias = torch.einsum('sbn,sbn->sb', A_neg, b_neg.expand_as(A_neg))
        return A_new, bias

    @staticmethod
    def backward(ctx, gA_out, gbias):
        A, d_pos, d_neg, b_neg = ctx.saved_tensors
        gbias = gbias.unsqueeze(-1)
        pos = (A >= 0).to(gA_out.dtype)
        neg = 1. - pos
        pg = pos * gA_out
        ng = neg * gA_out
        gd_pos = A * pg
        gd_neg = A * ng
        gb_neg = A * (neg * gbias)
        gA = d_pos * pg + d_neg * ng + b_neg * (neg * gbias)

        def _reduce(g, ref):
            while g.dim() > ref.dim():
                g = g.sum(0)
            for i, (a, b) in enumerate(zip(g.shape, ref.shape)):
                if a != b:
                    g = g.sum(i, keepdim=True)
            return g
        return gA, _reduce(gd_pos, d_pos), _reduce(gd_neg, d_neg), _reduce(gb_neg, b_neg)


# ----------------------------------------------------------------------------------------------
# one backward pass
# ----------------------------------------------------------------------------------------------

def _conv_output_padding(node, in_shape):
    """AL/operators/convolution.py:66-75."""
    w = node['weight']
    s, p, d = node['stride'], node['padding'], node['dilation']
    out_shape = node['shape']
    op0 = in_shape[1] - (out_shape[1] - 1) * s[0] + 2 * p[0] - 1 - (w.shape[2] - 1) * d[0]
    op1 = in_shape[2] - (out_shape[2] - 1) * s[1] + 2 * p[1] - 1 - (w.shape[3] - 1) * d[0]
    return (int(op0), int(op1))


def bn_affine(node):
    """AL/operators/normalization.py:117-118."""
    tmp_weight = node['weight'] / torch.sqrt(node['var'] + node['eps'])
    tmp_bias = node['bias'] - node['mean'] / torch.sqrt(node['var'] + node['eps']) * node['weight']
    return tmp_weight, tmp_bias


def crown_pass(nodes: List[dict], C: torch.Tensor, x_L: torch.Tensor, x_U: torch.Tensor,
               lower: Dict[int, torch.Tensor], upper: Dict[int, torch.Tensor],
               alpha: Optional[Dict[int, torch.Tensor]] = None,
               alpha_index: Optional[Dict[int, Optional[torch.Tensor]]] = None,
               beta: Optional[Dict[int, dict]] = None,
               sshape=None):
    """One backward CROWN pass from the output node, lower bound only.

    C        [Bd,S,n_out]            x_L,x_U [Bd,*in_shape]
    lower/upper[k]  [Bd,*shape_k]    keyed by the index of the PRE-activation node
    alpha[r] [S1,Bd,n_alpha]         keyed by the index of the activation node (plane 0 of the
                                     reference's [2,S1,Bd,n_alpha]); None => CROWN-adaptive slope.
                                     For sigmoid/tanh nodes: the FULL parameter [8,S1,Bd,*shape]
                                     (clipped in place by the pass, AL/operators/tanh.py:191-198)
    alpha_index[r] int64 [n_alpha] (flattened neuron ids) or None for dense alpha
    beta[k]  {'val','loc','sign','bias'} each [Bd,J], keyed by pre-activation node index
    returns  lb [Bd,S], lA {r: [S,Bd,*shape]}
    """
    Bd, S = C.shape[0], C.shape[1]
    n_nodes = len(nodes)
    out = n_nodes - 1
    A: List[Optional[torch.Tensor]] = [None] * n_nodes
    A[out] = C.transpose(0, 1).reshape(S, Bd, *nodes[out]['shape'])
    lb = torch.zeros(S, Bd, dtype=C.dtype)
    lAs = {}

    def _acc(i, val):
        A[i] = val if A[i] is None else A[i] + val

    for idx in range(n_nodes - 1, 0, -1):
        node = nodes[idx]
        a = A[idx]
        if a is None:
            continue
        op = node['op']
        # beta is injected on the pre-activation node's A before it propagates (backward_bound.py:222-227)
        if beta is not None and idx in beta and idx != out:
            bt = beta[idx]
            vals = (bt['val'] * bt['sign']).unsqueeze(0).expand(S, -1, -1)
            loc = bt['loc'].unsqueeze(0).expand(S, -1, -1)
            a = a.reshape(S, Bd, -1).scatter_add(2, loc, -vals).view(a.shape)
            if bt.get('bias') is not None:
                lb = lb + (vals * bt['bias'].unsqueeze(0)).sum(-1)
        if op == 'linear':
            _acc(node['in'][0], a.matmul(node['weight']))
            if node.get('bias') is not None:
                lb = lb + a.matmul(node['bias'])
        elif op == 'conv2d':
            src = nodes[node['in'][0]]
            shape = a.shape
            nxt = F.conv_transpose2d(a.reshape(S * Bd, *shape[2:]), node['weight'], None,
                                     stride=node['stride'], padding=node['padding'],
                                     dilation=node['dilation'], groups=node['groups'],
                                     output_padding=_conv_output_padding(node, src['shape']))
            _acc(node['in'][0], nxt.view(S, Bd, *nxt.shape[1:]))
            if node.get('bias') is not None:
                lb = lb + torch.einsum('sbchw,c->sb', a, node['bias'])
        elif op == 'batchnorm2d':
            w, b = bn_affine(node)
            _acc(node['in'][0], a * w.view(1, 1, -1, 1, 1))
            lb = lb + (a.sum((3, 4)) * b).sum(2)
        elif op == 'add':
            _acc(node['in'][0], a)
            _acc(node['in'][1], a)
        elif op == 'sub':
            _acc(node['in'][0], a)
            _acc(node['in'][1], -a)
        elif op == 'flatten':
            src = nodes[node['in'][0]]
            _acc(node['in'][0], a.reshape(S, Bd, *src['shape']))
        elif op == 'addconst':
            # unperturbed operand of Add/Sub: AL/backward_bound.py:712-721, AL/operators/base.py:320-341
            _acc(node['in'][0], a)
            lb = lb + torch.einsum('sb...,...->sb', a, node['value'])
        elif op == 'relu':
            k = node['in'][0]
            l, u = lower[k], upper[k]
            upper_d, upper_b = relu_upper_line(l, u)
            n = l[0].numel()
            if alpha is not None and idx in alpha:
                ai = alpha_index.get(idx) if alpha_index is not None else None
                a_full = reconstruct_full_alpha(alpha[idx], n, ai).view(-1, Bd, *l.shape[1:])
                lower_d = relu_lower_slope(l, u, a_full, upper_d)
            else:
                lower_d = relu_lower_slope(l, u, None, upper_d)
            lAs[idx] = a
            shp = a.shape
            new_A, bias = _SignSplit.apply(a.reshape(S, Bd, n), lower_d.reshape(-1, Bd, n),
                                           upper_d.reshape(1, Bd, n), upper_b.reshape(1, Bd, n))
            _acc(k, new_A.view(shp))
            lb = lb + bias
        elif op in ('sigmoid', 'tanh'):
            if sshape is None:
                from oracle import sshape_oracle as sshape
            k = node['in'][0]
            lAs[idx] = a
            new_A, bias = sshape.backward(op, a, lower[k], upper[k],
                                          alpha.get(idx) if alpha is not None else None)
            _acc(k, new_A)
            lb = lb + bias
        else:
            raise NotImplementedError(op)

    a0 = A[0].reshape(S, Bd, -1).transpose(0, 1)              # [Bd,S,n_in]
    x_ub = x_U.reshape(Bd, -1, 1)
    x_lb = x_L.reshape(Bd, -1, 1)
    center = (x_ub + x_lb) / 2.0
    diff = (x_ub - x_lb) / 2.0
    bound = a0.matmul(center) - a0.abs().matmul(diff)         # perturbations.py:170-171, sign=-1
    lb = lb.transpose(0, 1) + bound.squeeze(-1)
    return lb, lAs


# ----------------------------------------------------------------------------------------------
# the alpha/beta-CROWN optimisation loop
# ----------------------------------------------------------------------------------------------

def optimize(nodes, C, x_L, x_U, lower, upper, alpha, alpha_index, beta, rhs,
             iteration=20, lr_alpha=0.1, lr_beta=0.1, lr_decay=0.98,
             early_stop_patience=10, start_save_best=0.5, enable_beta=True, sshape=None,
             trace: Optional[list] = None):
    """AL/optimized_bounds.py:255-629 with bound_side='lower', keep_best=True,
    fix_interm_bounds=True, stop_criterion_batch_any(rhs) (AL/utils.py:87-93).

    alpha[r]: the reference's FULL parameter tensor [2,S1,Bd,n_alpha] (both planes are Adam
    parameters; only plane 0 is used by a lower-bound pass).
    beta[k]['val'] [Bd,J] is the other parameter group.
    Returns dict(lb=best lower bounds [Bd,S], alpha=best alphas, beta_val=best betas,
                 lA=lA of the LAST executed pass, n_iter=iterations executed).
    """
    params_a = {r: a.detach().clone().requires_grad_() for r, a in alpha.items()}
    use_beta = enable_beta and beta is not None and len(beta) > 0
    params_b = {}
    if use_beta:
        params_b = {k: b['val'].detach().clone().requires_grad_() for k, b in beta.items()}
    groups = [{'params': list(params_a.values()), 'lr': lr_alpha}]
    if use_beta:
        groups.append({'params': list(params_b.values()), 'lr': lr_beta})
    opt = torch.optim.Adam(groups)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, lr_decay)
    best_alpha = {r: a.detach().clone() for r, a in params_a.items()}
    best_beta = {k: b.detach().clone() for k, b in params_b.items()}

    def red(t):  # loss_reduction_func = sum over specs, applied when S != 1
        return t.sum(1, keepdim=True) if t.shape[1] != 1 else t

    patience = 0
    best_l = None
    ret_0 = None
    lAs = None
    n_iter = 0
    for i in range(iteration):
        need_grad = i != iteration - 1
        cur_beta = None
        if beta is not None and len(beta) > 0 and enable_beta:
            cur_beta = {k: dict(b, val=params_b[k]) for k, b in beta.items()}
        with torch.enable_grad() if need_grad else torch.no_grad():
            lb, lAs = crown_pass(nodes, C, x_L, x_U, lower, upper,
                                 {r: (a[0] if nodes[r]['op'] == 'relu' else a) for r, a in params_a.items()},
                                 alpha_index, cur_beta, sshape=sshape)
        n_iter = i + 1
        full = lb.detach()
        if i == 0:
            best_l = torch.full_like(full, float('-inf'))
            best_ret = full.clone()
            ret_0 = full.clone()
        stop = (full > rhs).any(dim=1, keepdim=True)
        l = red(lb)
        loss = ((-1 * l) * stop.logical_not()).sum()
        stop_final = bool(stop.all())
        with torch.no_grad():
            mask = (red(full) > red(best_l)).view(-1)
            need_update = bool(mask.any())
            if need_update:
                idx = mask.nonzero(as_tuple=True)[0]
                best_l[idx] = torch.max(full[idx], best_l[idx])
                best_ret[idx] = torch.max(full[idx], best_ret[idx])
            patience = 0 if need_update else patience + 1
            if (i < 1 or i > int(iteration * start_save_best) or stop_final
                    or patience == early_stop_patience):
                mask0 = (red(full) > red(ret_0)).view(-1)
                if mask0.any():
                    idx = mask0.nonzero(as_tuple=True)[0]
                    ret_0[idx] = full[idx]
                    for r in params_a:
                        best_alpha[r][:, :, idx] = params_a[r].detach()[:, :, idx]
                    for k in params_b:
                        best_beta[k][idx] = params_b[k].detach()[idx]
                else:
                    # Reference quirk (optimized_bounds.py:487-491): _get_idx_mask returns idx=None
                    # when nothing improved and `ret_0[None] = full_ret_l[None]` then overwrites
                    # ret_0 for EVERY domain with the current (not better) bounds; no snapshot.
                    ret_0 = full.clone()
        if trace is not None:
            trace.append({'lb': full.clone(),
                          'alpha': {r: a.detach().clone() for r, a in params_a.items()},
                          'beta': {k: b.detach().clone() for k, b in params_b.items()}})
        if stop_final:
            break
        if patience > early_stop_patience:
            break
        opt.zero_grad(set_to_none=True)
        if need_grad:
            loss.backward()
            if trace is not None:
                trace[-1]['grad_alpha'] = {r: a.grad.detach().clone() for r, a in params_a.items()}
                trace[-1]['grad_beta'] = {k: b.grad.detach().clone() for k, b in params_b.items()}
            opt.step()
            sched.step()
        with torch.no_grad():
            for b in params_b.values():
                b.data = (b >= 0) * b.data
            for r, a in params_a.items():
                if nodes[r]['op'] == 'relu':          # clip_alpha: AL/operators/relu.py:104-108; a no-op
                    a.data = torch.clamp(a.data, 0., 1.)   # for S-shapes (activation_base.py:203-204)
    return {'lb': best_ret, 'alpha': best_alpha, 'beta_val': best_beta,
            'lA': {r: v.detach() for r, v in lAs.items()}, 'n_iter': n_iter}


# ----------------------------------------------------------------------------------------------
# helpers shared by tests / bench (synthetic sub-domain batches)
# ----------------------------------------------------------------------------------------------

def forward(nodes: List[dict], x: torch.Tensor):
    """Concrete forward of the node list (used for IBP-free sanity checks and spec construction)."""
    vals = [None] * len(nodes)
    vals[0] = x
    for i, nd in enumerate(nodes):
        op = nd['op']
        if op == 'input':
            continue
        a = vals[nd['in'][0]]
        if op == 'linear':
            vals[i] = F.linear(a, nd['weight'], nd.get('bias'))
        elif op == 'conv2d':
            vals[i] = F.conv2d(a, nd['weight'], nd.get('bias'), nd['stride'], nd['padding'],
                               nd['dilation'], nd['groups'])
        elif op == 'batchnorm2d':
            w, b = bn_affine(nd)
            vals[i] = a * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        elif op == 'add':
            vals[i] = a + vals[nd['in'][1]]
        elif op == 'sub':
            vals[i] = a - vals[nd['in'][1]]
        elif op == 'flatten':
            vals[i] = a.flatten(1)
        elif op == 'addconst':
            vals[i] = a + nd['value']
        elif op == 'relu':
            vals[i] = F.relu(a)
        elif op == 'sigmoid':
            vals[i] = torch.sigmoid(a)
        elif op == 'tanh':
            vals[i] = torch.tanh(a)
        else:
            raise NotImplementedError(op)
    return vals


def interval_bounds(nodes: List[dict], x_L: torch.Tensor, x_U: torch.Tensor):
    """Plain IBP over the node list -> {preact_idx: (l,u)}; only used to fabricate valid
    (sound, loose) intermediate bounds for synthetic parity inputs."""
    lo = [None] * len(nodes)
    hi = [None] * len(nodes)
    lo[0], hi[0] = x_L, x_U
    pre = {}
    for i, nd in enumerate(nodes):
        op = nd['op']
        if op == 'input':
            continue
        a_l, a_u = lo[nd['in'][0]], hi[nd['in'][0]]
        if op in ('linear', 'conv2d', 'batchnorm2d'):
            c, r = (a_l + a_u) / 2, (a_u - a_l) / 2
            if op == 'linear':
                w = nd['weight']
                cc = F.linear(c, w, nd.get('bias'))
                rr = F.linear(r, w.abs())
            elif op == 'conv2d':
                w = nd['weight']
                cc = F.conv2d(c, w, nd.get('bias'), nd['stride'], nd['padding'], nd['dilation'], nd['groups'])
                rr = F.conv2d(r, w.abs(), None, nd['stride'], nd['padding'], nd['dilation'], nd['groups'])
            else:
                w, b = bn_affine(nd)
                cc = c * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
                rr = r * w.abs().view(1, -1, 1, 1)
            lo[i], hi[i] = cc - rr, cc + rr
        elif op == 'add':
            lo[i], hi[i] = a_l + lo[nd['in'][1]], a_u + hi[nd['in'][1]]
        elif op == 'sub':
            lo[i], hi[i] = a_l - hi[nd['in'][1]], a_u - lo[nd['in'][1]]
        elif op == 'flatten':
            lo[i], hi[i] = a_l.flatten(1), a_u.flatten(1)
        elif op == 'addconst':
            lo[i], hi[i] = a_l + nd['value'], a_u + nd['value']
        elif op in ('relu', 'sigmoid', 'tanh'):
            pre[nd['in'][0]] = (a_l, a_u)
            f = {'relu': F.relu, 'sigmoid': torch.sigmoid, 'tanh': torch.tanh}[op]
            lo[i], hi[i] = f(a_l), f(a_u)
        else:
            raise NotImplementedError(op)
    return pre
