"""Host-side mirror of the `BoundedModule` surface that NeuralSAT's abstractor / heuristics / verifier
consume, backed by the CUDA library through `capi.Plan`.

Reference interface mirrored here (paths under /root/reference/neuralsat-pt201, AL = auto_LiRPA):
  BoundedModule.__init__ ................ AL/bound_general.py:37-150   (graph construction)
  BoundedModule.compute_bounds .......... AL/bound_general.py:921-1179 (the hot path entry)
  BoundedModule.set_bound_opts .......... AL/bound_general.py:224-230  (nested dict update)
  BoundedModule.get_split_nodes ......... AL/beta_crown.py:45-70
  BoundedModule.reset_beta / SparseBeta . AL/beta_crown.py:11-42, :101-114
  BoundedModule.get_splittable_activations, nodes(), __getitem__, final_name, input_name,
  root_names, relus, perturbed_optimizable_activations, layers_requiring_bounds
  BoundedTensor / PerturbationLpNorm .... AL/bounded_tensor.py, AL/perturbations.py:101-183
  stop_criterion_batch_any .............. AL/utils.py:87-93

Same names, argument meaning and error behaviour as the reference for the calls the BaB loop makes
(`NetworkAbstractor._forward_hidden`, NS/abstractor/abstractor.py:244-344): fixed intermediate bounds
(`interm_bounds` for every split node), `method in {'backward','crown-optimized'}`, lower bound only.
Everything else (root bounds without interm_bounds, forward mode, IBP, Gurobi members) is outside the
hot path and raises NotImplementedError.  There is no CPU path: a CUDA device and libcrown_b200.so
are required, and allocation failures surface as RuntimeError('CUDA out of memory. ...') exactly as
the reference's batch-halving logic expects (NS/util/misc/torch_cuda_memory.py:58-61).
"""
from __future__ import annotations

import copy
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import capi
from .graph import ACTIVATIONS, nodes_to, trace_module


# ----------------------------------------------------------------------------------------------
# small value types mirrored from auto_LiRPA
# ----------------------------------------------------------------------------------------------
class PerturbationLpNorm:
    """L-inf box with explicit bounds (AL/perturbations.py:101-183); only norm=inf is on the path."""

    def __init__(self, eps=0, norm=float('inf'), x_L=None, x_U=None):
        if norm != float('inf'):
            raise NotImplementedError('only L-inf perturbations are on the hot path')
        self.eps, self.norm, self.x_L, self.x_U = eps, norm, x_L, x_U


class BoundedTensor(torch.Tensor):
    """A tensor carrying its perturbation (AL/bounded_tensor.py)."""

    @staticmethod
    def __new__(cls, x, ptb=None, *args, **kwargs):
        return torch.Tensor._make_subclass(cls, x.data if isinstance(x, torch.Tensor) else torch.as_tensor(x), False)

    def __init__(self, x, ptb=None):
        self.ptb = ptb

    def to(self, *args, **kwargs):
        t = super().to(*args, **kwargs)
        ptb = self.ptb
        if ptb is not None:
            ptb = PerturbationLpNorm(ptb.eps, ptb.norm,
                                     None if ptb.x_L is None else ptb.x_L.to(*args, **kwargs),
                                     None if ptb.x_U is None else ptb.x_U.to(*args, **kwargs))
        return BoundedTensor(t.as_subclass(torch.Tensor), ptb)


class _StopAny:
    """`stop_criterion_batch_any(rhs)` (AL/utils.py:87-93) as an object, so that the threshold can be
    handed to the device loop instead of being called back per iteration."""

    def __init__(self, threshold):
        self.threshold = threshold

    def __call__(self, x):
        return (x > self.threshold).any(dim=1, keepdim=True)


def stop_criterion_batch_any(threshold):
    return _StopAny(threshold)


def _threshold_of(func):
    """rhs captured by a stop criterion: ours, or the reference's closure (AL/utils.py:87-93)."""
    if func is None:
        return None
    if hasattr(func, 'threshold'):
        return func.threshold
    for cell in getattr(func, '__closure__', None) or ():
        v = cell.cell_contents
        if isinstance(v, torch.Tensor):
            return v
    return None


def _ragged_fill(dst: torch.Tensor, rows) -> None:
    """dst[i, :len(rows[i])] = rows[i] for every i, with ONE concatenation and ONE indexed store instead of a tensor
    operation per domain (the reference loops, AL/beta_crown.py:25-42).  rows: sequence of 1-d tensors / lists / None."""
    lens = [0 if r is None else (r.shape[0] if isinstance(r, torch.Tensor) else len(r)) for r in rows]
    total = sum(lens)
    if total == 0:
        return
    parts = [r if isinstance(r, torch.Tensor) else torch.as_tensor(r) for r, n in zip(rows, lens) if n > 0]
    flat = torch.cat([q.reshape(-1) if q.dim() != 1 else q for q in parts]).to(device='cpu', dtype=dst.dtype)
    lens_t = torch.tensor(lens)
    row_idx = torch.repeat_interleave(torch.arange(len(rows)), lens_t)
    starts = torch.cumsum(lens_t, 0) - lens_t
    col_idx = torch.arange(total) - torch.repeat_interleave(starts, lens_t)
    dst[row_idx, col_idx] = flat


class SparseBeta:
    """AL/beta_crown.py:11-42: per split node, `val/loc/sign(/bias)` of shape [Bd, Jmax], zero padded
    (padded entries have sign 0)."""

    def __init__(self, shape, bias=False, betas=None, device='cpu'):
        self.device = device
        val = torch.zeros(shape)
        self.loc = torch.zeros(shape, dtype=torch.long)
        self.sign = torch.zeros(shape)
        self.bias = torch.zeros(shape) if bias else None
        if betas:
            _ragged_fill(val, betas)
        self.val = val.to(device, non_blocking=True)

    def apply_splits(self, history, key):
        """Fill loc/sign(/bias) from the per-domain split histories `history[bi][key] = (loc, sign, point)`."""
        loc, sign, bias = self.loc.cpu(), self.sign.cpu(), None if self.bias is None else self.bias.cpu()
        _ragged_fill(loc, [h[key][0] for h in history])
        _ragged_fill(sign, [h[key][1] for h in history])
        if bias is not None:
            _ragged_fill(bias, [h[key][2] for h in history])
        self.loc = loc.to(self.device, non_blocking=True)
        self.sign = sign.to(self.device, non_blocking=True)
        if bias is not None:
            self.bias = bias.to(self.device, non_blocking=True)


class _Node:
    """One graph node as seen through the reference API (`Bound` attributes the callers read)."""

    def __init__(self, index, desc):
        self.index = index
        self.name = desc['name']
        self.op = desc['op']
        self.output_shape = (1, *desc['shape'])
        self.inputs: List['_Node'] = []
        self.output_name: List[str] = []
        self.lower = None
        self.upper = None
        self.perturbed = True
        # pre-activation nodes
        self.sparse_betas = None
        # activation nodes
        self.alpha: Dict[str, torch.Tensor] = {}
        self.alpha_indices = None          # tuple of index tensors (OP/relu.py:330-332) or None = dense
        self.lA = None
        self._alpha_pos = None             # cached int32 neuron -> column map

    def get_split_point(self):
        """0.0 for ReLU (OP/relu.py:327-328), None for S-shaped activations (OP/tanh.py:306-307)."""
        return 0.0 if self.op == 'relu' else None

    def __repr__(self):
        return f'Bound{self.op.capitalize()}(name={self.name})'


_DEFAULT_OPT = {
    'enable_alpha_crown': True, 'enable_beta_crown': False, 'iteration': 20, 'lr_alpha': 0.5,
    'lr_beta': 0.05, 'lr_decay': 0.98, 'early_stop_patience': 10, 'start_save_best': 0.5,
    'fix_interm_bounds': True, 'stop_criterion_func': None, 'use_shared_alpha': False,
    'init_alpha': False, 'keep_best': True,
}


class BoundedModule(nn.Module):
    """Drop-in for the calls NeuralSAT's BaB loop makes on `auto_LiRPA.BoundedModule`."""

    def __init__(self, model: nn.Module, global_input, bound_opts=None, device='cuda', verbose=False,
                 custom_ops=None):
        super().__init__()
        if custom_ops:
            raise NotImplementedError('custom operators are outside the hot path (SURVEY.md 8b)')
        dev = torch.device(device)
        if dev.type != 'cuda':
            raise RuntimeError('neuralsat_b200.BoundedModule needs a CUDA device: there is no CPU path')
        self.device = dev
        self.ori_model = model
        shape = tuple(global_input.shape) if isinstance(global_input, torch.Tensor) else tuple(global_input)
        self.graph = trace_module(copy.deepcopy(model).to('cpu'), shape)
        self.plan = capi.Plan(nodes_to(self.graph, dev))
        self.bound_opts = {'conv_mode': 'matrix', 'crown_batch_size': 10 ** 9,
                           'optimize_bound_args': dict(_DEFAULT_OPT)}
        if bound_opts:
            self.set_bound_opts(bound_opts)
        self._nodes = [_Node(i, d) for i, d in enumerate(self.graph)]
        for nd, d in zip(self._nodes, self.graph):
            for j in d.get('in', []):
                nd.inputs.append(self._nodes[j])
                self._nodes[j].output_name.append(nd.name)
        self._by_name = {n.name: n for n in self._nodes}
        self.input_name = [self._nodes[0].name]
        self.root_names = [self._nodes[0].name]
        self.final_name = self._nodes[-1].name
        self.final_node_name = self.final_name
        self.relus = [self._nodes[i] for i in self.plan.act_nodes if self._nodes[i].op == 'relu']
        self.perturbed_optimizable_activations = [self._nodes[i] for i in self.plan.act_nodes]
        self.layers_requiring_bounds = [self._nodes[i] for i in self.plan.pre_nodes]
        self.split_nodes: List[_Node] = []
        self.split_activations: Dict[str, list] = {}
        self.get_split_nodes()
        self.last_n_iter = 0
        self._graph_dev = None            # node list on the device (root-phase interval arithmetic)
        self._prefix_plans: Dict[int, capi.Plan] = {}
        self.root_workspace_bytes = 2 << 30       # spec rows of an intermediate-bound pass are chunked to this workspace

    # ---- graph accessors (AL/bound_general.py:232-262) -------------------------------------------
    def nodes(self):
        return list(self._nodes)

    def __getitem__(self, name):
        return self._by_name[name]

    def final_node(self):
        return self._nodes[-1]

    def get_splittable_activations(self):
        return list(self.perturbed_optimizable_activations)

    def get_split_nodes(self, input_split=False):
        """AL/beta_crown.py:45-70."""
        self.split_nodes, self.split_activations = [], {}
        for act in self.perturbed_optimizable_activations:
            pre = act.inputs[0]
            if pre not in self.split_nodes:
                self.split_nodes.append(pre)
                self.split_activations[pre.name] = []
            self.split_activations[pre.name].append((act, 0))
        if input_split:
            root = self._nodes[0]
            if root not in self.split_nodes:
                self.split_nodes.append(root)
                self.split_activations[root.name] = []
        return self.split_nodes, self.split_activations

    def forward(self, x):
        return self.ori_model(x)

    # ---- hand-over from a reference BoundedModule (variant A) ----------------------
    def name_map(self, reference_net) -> Dict[str, str]:
        """reference node name -> node name here, for the names the BaB loop uses as dictionary keys:
        activations (`relus` / `perturbed_optimizable_activations`), pre-activation (`split_nodes`)
        and the final node.  Both tracers keep program order (AL/bound_general.py:286-318)."""
        ref_acts = list(reference_net.perturbed_optimizable_activations)
        if len(ref_acts) != len(self.perturbed_optimizable_activations):
            raise ValueError('the two graphs have a different number of activations')
        m = {reference_net.final_name: self.final_name}
        for ra, a in zip(ref_acts, self.perturbed_optimizable_activations):
            m[ra.name] = a.name
            m[ra.inputs[0].name] = a.inputs[0].name
        return m

    def adopt_root_state(self, reference_net) -> Dict[str, str]:
        """Take over what `initialize()` left on the reference module: alpha tensors of the final start
        node, their sparse-feature `alpha_indices` (OP/relu.py:330-332) and the intermediate bounds."""
        m = self.name_map(reference_net)
        for ra, a in zip(reference_net.perturbed_optimizable_activations, self.perturbed_optimizable_activations):
            al = getattr(ra, 'alpha', None) or {}
            if reference_net.final_name in al:
                a.alpha = {self.final_name: al[reference_net.final_name].detach().to(self.device, torch.float32).contiguous()}
            idx = getattr(ra, 'alpha_indices', None)
            a.alpha_indices = None if idx is None else tuple(i.to(self.device) for i in idx)
            a._alpha_pos = None
            pre_r, pre = ra.inputs[0], a.inputs[0]
            if getattr(pre_r, 'lower', None) is not None:
                pre.lower = pre_r.lower.detach().to(self.device)
                pre.upper = pre_r.upper.detach().to(self.device)
        return m

    # ---- options (AL/bound_general.py:224-230: nested dict UPDATE) ---------------------------------
    def set_bound_opts(self, new_opts):
        for k, v in new_opts.items():
            if isinstance(v, dict) and isinstance(self.bound_opts.get(k), dict):
                self.bound_opts[k].update(v)
            else:
                self.bound_opts[k] = v

    # ---- beta (AL/beta_crown.py:101-114) ----------------------------------------------------------
    def reset_beta(self, node, shape, betas, bias=False, start_nodes=None):
        node.sparse_betas = [SparseBeta(shape=shape, betas=betas, device=self.device, bias=bool(bias))]

    # ---- alpha ------------------------------------------------------------------------------------
    def init_alpha(self, x, share_alphas=False, method='backward', c=None, interm_bounds=None, **kw):
        """Creates `act.alpha[final_name]` = CROWN-adaptive slope (OP/relu.py:101-103, :244-246) from
        the given intermediate bounds.  Computing those bounds from scratch (the reference's
        init_alpha runs a full CROWN, AL/optimized_bounds.py:632-723) is a root-bounds task."""
        lb = None
        if interm_bounds is None:
            # the reference's init_alpha first runs one plain CROWN with every intermediate layer bounded
            # (AL/optimized_bounds.py:661); that leaves node.lower / node.upper on every split node
            if x is None or c is None:
                raise ValueError('init_alpha without interm_bounds needs x and c')
            lb, _ = self.compute_bounds(x=x, C=c, method='backward', bound_upper=False)
            interm_bounds = {n.name: [n.lower, n.upper] for n in self.layers_requiring_bounds}
        S1 = 1 if c is None else int(c.shape[1])
        sparse = bool(self.bound_opts.get('sparse_features_alpha', True))
        min_sparsity = float(self.bound_opts.get('minimum_sparsity', 0.9))
        for act in self.perturbed_optimizable_activations:
            l, u = interm_bounds[act.inputs[0].name]
            l, u = l.to(self.device), u.to(self.device)
            if act.op != 'relu':
                # [8,S1,Bd,*shape] tangent points (OP/tanh.py:54-63): middle point, then the table points
                from .sshape_tables import lookup_points
                d_lower, d_upper = lookup_points(act.op, l, u)
                a = torch.empty(8, S1, *l.shape, device=self.device)
                a[:4] = (l + u) / 2
                a[4:6] = d_lower
                a[6:8] = d_upper
                act.alpha = {self.final_name: a}
                act.alpha_indices = None
                act._alpha_pos = None
                continue
            lb_r, ub_r = l.clamp(max=0), u.clamp(min=0)
            ub_r = torch.max(ub_r, lb_r + 1e-8)
            init = ((ub_r / (ub_r - lb_r)) > 0.5).to(l.dtype)
            act.alpha_indices = None
            act._alpha_pos = None
            if lb is not None and sparse:
                # sparse-feature alpha (OP/relu.py:37-77): slopes only for neurons unstable in SOME batch element,
                # unless more than `minimum_sparsity` of the layer is
                unstable = torch.logical_and(l < 0, u > 0).any(dim=0).flatten().nonzero().flatten()
                if unstable.numel() <= min_sparsity * l[0].numel():
                    act.alpha_indices = (unstable,)
                    init = init.flatten(1)[:, unstable]
            act.alpha = {self.final_name: init.unsqueeze(0).unsqueeze(0).repeat(2, S1, *([1] * init.dim())).contiguous()}
        if lb is not None:
            aux = {n.name: [n.lower.detach().clone(), n.upper.detach().clone()] for n in self.layers_requiring_bounds}
            return lb, None, aux

    # ---- root phase: intermediate bounds of every layer -------------------------
    def _dev_graph(self):
        if self._graph_dev is None:
            self._graph_dev = nodes_to(self.graph, self.device)
        return self._graph_dev

    def _interval_of(self, idx, known, memo):
        """Interval bounds of node `idx` from the nearest nodes with known bounds (`known`: the input box and the
        pre-activation nodes bounded so far), the arithmetic of the operators' `interval_propagate`
        (AL/interval_bound.py:16-145, OP/linear.py:418-470 for L-inf: centre / deviation form)."""
        if idx in known:
            return known[idx]
        if idx in memo:
            return memo[idx]
        import torch.nn.functional as F
        nd = self._dev_graph()[idx]
        op = nd['op']
        lo, hi = self._interval_of(nd['in'][0], known, memo)
        if op in ('linear', 'conv2d', 'batchnorm2d'):
            mid, diff = (lo + hi) / 2.0, (hi - lo) / 2.0
            if op == 'linear':
                center = F.linear(mid, nd['weight'], nd.get('bias'))
                dev = F.linear(diff, nd['weight'].abs())
            elif op == 'conv2d':
                args = (nd['stride'], nd['padding'], nd['dilation'], nd['groups'])
                center = F.conv2d(mid, nd['weight'], nd.get('bias'), *args)
                dev = F.conv2d(diff, nd['weight'].abs(), None, *args)
            else:
                w = nd['weight'] / torch.sqrt(nd['var'] + nd['eps'])
                b = nd['bias'] - nd['mean'] / torch.sqrt(nd['var'] + nd['eps']) * nd['weight']
                center = mid * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
                dev = diff * w.abs().view(1, -1, 1, 1)
            out = (center - dev, center + dev)
        elif op == 'relu':
            out = (lo.clamp(min=0), hi.clamp(min=0))
        elif op == 'sigmoid':
            out = (torch.sigmoid(lo), torch.sigmoid(hi))
        elif op == 'tanh':
            out = (torch.tanh(lo), torch.tanh(hi))
        elif op == 'add':
            l2, h2 = self._interval_of(nd['in'][1], known, memo)
            out = (lo + l2, hi + h2)
        elif op == 'sub':
            l2, h2 = self._interval_of(nd['in'][1], known, memo)
            out = (lo - h2, hi - l2)
        elif op == 'flatten':
            out = (lo.flatten(1), hi.flatten(1))
        elif op == 'addconst':
            out = (lo + nd['value'], hi + nd['value'])
        else:
            raise NotImplementedError(op)
        memo[idx] = out
        return out

    def _prefix_plan(self, idx):
        """Plan of the sub-network that ends in node `idx` (a pre-activation node): its backward pass with an
        identity specification bounds that layer (compute_intermediate_bounds, AL/bound_general.py:782-903)."""
        if idx not in self._prefix_plans:
            self._prefix_plans[idx] = capi.Plan(self._dev_graph()[:idx + 1])
        return self._prefix_plans[idx]

    def _bound_layer(self, idx, x_L, x_U, known, select):
        """CROWN lower and upper bounds of the neurons `select` (flattened ids, None = all) of node `idx`:
        one pass over the prefix plan with spec rows [+e_j ; -e_j], upper = -lower(-e_j)."""
        plan = self._prefix_plan(idx)
        Bd = int(x_L.shape[0])
        n = 1
        for d in self.graph[idx]['shape']:
            n *= int(d)
        sel = torch.arange(n, device=self.device) if select is None else select
        lower = [known[p][0] for p in plan.pre_nodes]
        upper = [known[p][1] for p in plan.pre_nodes]
        per_row = 4 * sum(int(torch.tensor(nd['shape']).prod()) for nd in self.graph[:idx + 1]) + 64
        chunk = max(1, int(self.root_workspace_bytes // (2 * Bd * per_row)))
        lo = torch.empty(Bd, sel.numel(), device=self.device)
        hi = torch.empty(Bd, sel.numel(), device=self.device)
        for s0 in range(0, sel.numel(), chunk):
            ids = sel[s0:s0 + chunk]
            S = ids.numel()
            C = torch.zeros(Bd, 2 * S, n, device=self.device)
            ar = torch.arange(S, device=self.device)
            C[:, ar, ids] = 1.0
            C[:, ar + S, ids] = -1.0
            lb, _ = plan.crown_pass(C, x_L, x_U, lower, upper, None, None, None, want_lA=False)
            lo[:, s0:s0 + S] = lb[:, :S]
            hi[:, s0:s0 + S] = -lb[:, S:]
        return lo, hi

    def _intermediate_bounds(self, x_L, x_U, reference_bounds=None):
        """Bounds of every split node, layer by layer (check_prior_bounds, AL/bound_general.py:739-780): the first
        layer by interval arithmetic (check_IBP_first_linear), later layers by CROWN from that layer - only for the
        neurons that interval arithmetic leaves unstable in some batch element when those are at most
        `minimum_sparsity` of the layer (get_sparse_C / restore_sparse_bounds, AL/backward_bound.py:339-552)."""
        Bd = int(x_L.shape[0])
        known = {0: (x_L, x_U)}
        sparse = bool(self.bound_opts.get('sparse_intermediate_bounds', True))
        min_sparsity = float(self.bound_opts.get('minimum_sparsity', 0.9))
        for node in self.layers_requiring_bounds:
            idx = node.index
            memo = {}
            ibp_l, ibp_u = self._interval_of(idx, known, memo)
            first = len(self._ancestors_with_acts(idx)) == 0      # no relaxation upstream: interval arithmetic is exact
            act_op = self.graph[[a for a in self.plan.act_nodes if self.graph[a]['in'][0] == idx][0]]['op']
            if first:
                l, u = ibp_l, ibp_u
            else:
                select = None
                l, u = ibp_l.clone(), ibp_u.clone()
                if sparse and act_op == 'relu':
                    unstable = torch.logical_and(ibp_l < 0, ibp_u > 0).any(dim=0).flatten().nonzero().flatten()
                    total = ibp_l[0].numel()
                    if unstable.numel() == 0:
                        select = unstable
                    elif unstable.numel() <= min_sparsity * total:
                        select = unstable
                if select is None:
                    lo, hi = self._bound_layer(idx, x_L, x_U, known, None)
                    l, u = lo.view_as(ibp_l), hi.view_as(ibp_u)
                elif select.numel() > 0:
                    lo, hi = self._bound_layer(idx, x_L, x_U, known, select)
                    l.view(Bd, -1)[:, select] = lo
                    u.view(Bd, -1)[:, select] = hi
            if reference_bounds and node.name in reference_bounds:
                rl, ru = reference_bounds[node.name]
                l = torch.max(rl.to(self.device), l)
                u = torch.min(ru.to(self.device), u)
            known[idx] = (l.contiguous(), u.contiguous())
            node.lower, node.upper = known[idx]
        return {n.name: [n.lower, n.upper] for n in self.layers_requiring_bounds}

    def _ancestors_with_acts(self, idx):
        """Pre-activation nodes strictly upstream of node `idx` (empty for the first layer)."""
        seen, stack, out = set(), [idx], []
        pres = set(self.plan.pre_nodes)
        while stack:
            i = stack.pop()
            for j in self.graph[i].get('in', []):
                if j in seen:
                    continue
                seen.add(j)
                if j in pres:
                    out.append(j)
                stack.append(j)
        return out

    # ---- the hot path -----------------------------------------------------------------------------
    def _alpha_args(self, Bd, use_alpha):
        if not use_alpha:
            return None, None
        alphas, poss = [], []
        any_alpha = False
        for act in self.perturbed_optimizable_activations:
            a = act.alpha.get(self.final_name) if act.alpha else None
            if a is None:
                alphas.append(None)
                poss.append(None)
                continue
            if a.device != self.device or a.dtype != torch.float32 or not a.is_contiguous() or a.requires_grad:
                a = a.detach().to(self.device, torch.float32).contiguous()
                act.alpha[self.final_name] = a
            if a.shape[2] != Bd:
                raise ValueError(f'alpha of {act.name} has batch {a.shape[2]}, expected {Bd}')
            n = 1
            for s in act.output_shape[1:]:
                n *= int(s)
            pos = None
            if act.alpha_indices is not None:
                if act._alpha_pos is None:
                    idx = act.alpha_indices
                    if isinstance(idx, (tuple, list)):
                        flat, stride = torch.zeros_like(idx[0]), 1
                        for d, ix in zip(reversed(act.output_shape[1:]), reversed(idx)):
                            flat = flat + ix * stride
                            stride *= int(d)
                        idx = flat
                    act._alpha_pos = capi.alpha_pos_from_index(idx, n, self.device)
                pos = act._alpha_pos
            alphas.append(a)
            poss.append(pos)
            any_alpha = True
        return (alphas, poss) if any_alpha else (None, None)

    def _beta_args(self, Bd, use_beta):
        if not use_beta:
            return None
        out, any_beta = [], False
        for pre in self.layers_requiring_bounds:
            sb = pre.sparse_betas
            sb = sb[0] if isinstance(sb, list) and sb else None
            if sb is None or sb.val.shape[1] == 0:
                out.append(None)
                continue
            if sb.val.shape[0] != Bd:
                raise ValueError(f'beta of {pre.name} has batch {sb.val.shape[0]}, expected {Bd}')
            if sb.val.requires_grad or not sb.val.is_contiguous() or sb.val.device != self.device:
                sb.val = sb.val.detach().to(self.device).contiguous()
            out.append({'val': sb.val, 'loc': sb.loc.to(self.device), 'sign': sb.sign.to(self.device),
                        'bias': None if sb.bias is None else sb.bias.to(self.device)})
            any_beta = True
        return out if any_beta else None

    def compute_bounds(self, x=None, aux=None, C=None, method='backward', IBP=False, forward=False,
                       bound_lower=True, bound_upper=False, reuse_ibp=False, reuse_alpha=False,
                       return_A=False, needed_A_dict=None, final_node_name=None, average_A=False,
                       interm_bounds=None, reference_bounds=None, intermediate_constr=None,
                       alpha_idx=None, aux_reference_bounds=None, need_A_only=False, cutter=None,
                       decision_thresh=None, update_mask=None, **kwargs):
        """AL/bound_general.py:921-1179 for the calls of the BaB loop.  Returns `(lb [Bd,S], None)`."""
        if bound_upper or not bound_lower:
            raise NotImplementedError('the BaB loop only asks for lower bounds (SURVEY.md 8b)')
        if IBP or forward or return_A or cutter is not None or (final_node_name not in (None, self.final_name)):
            raise NotImplementedError('only backward bounds of the output node are on the hot path')
        m = str(method).lower()
        if m in ('backward', 'crown'):
            optimize = False
        elif m in ('crown-optimized', 'alpha-crown', 'crown_optimized', 'alpha-beta-crown'):
            optimize = True
        else:
            raise NotImplementedError(f'method {method!r} is outside the hot path')
        if x is None or C is None:
            raise ValueError('x=(BoundedTensor,) and C are required')
        xt = x[0] if isinstance(x, (tuple, list)) else x
        ptb = getattr(xt, 'ptb', None)
        if ptb is None or ptb.x_L is None or ptb.x_U is None:
            raise ValueError('x must be a BoundedTensor with explicit x_L/x_U (NS/abstractor/utils.py:24-27)')
        dev = self.device
        C = C.detach().to(dev, torch.float32).contiguous()
        Bd = int(C.shape[0])
        x_L = ptb.x_L.detach().to(dev, torch.float32).contiguous()
        x_U = ptb.x_U.detach().to(dev, torch.float32).contiguous()
        if interm_bounds is None or any(n.name not in interm_bounds for n in self.layers_requiring_bounds):
            # no fixed intermediate bounds (root of a verification, input-split regime): bound every layer first.
            # method='backward' reproduces the reference; for 'crown-optimized' the reference also re-tightens the
            # intermediate layers with their own slopes in every iteration (AL/optimized_bounds.py:345-376) - here
            # they are bounded once by CROWN and only the output node's slopes are optimised (sound, looser).
            given = dict(interm_bounds) if interm_bounds else {}
            have = all(n.lower is not None and n.lower.shape[0] == Bd for n in self.layers_requiring_bounds)
            if optimize and have and not given:
                interm_bounds = {n.name: [n.lower, n.upper] for n in self.layers_requiring_bounds}
            else:
                ref = dict(reference_bounds) if reference_bounds else {}
                ref.update(given)
                interm_bounds = self._intermediate_bounds(x_L, x_U, ref)
        lower, upper = [], []
        for n in self.layers_requiring_bounds:
            l, u = interm_bounds[n.name]
            # adopted by reference, detached (AL/bound_general.py:481-488)
            n.lower = l.detach().to(dev, torch.float32).contiguous()
            n.upper = u.detach().to(dev, torch.float32).contiguous()
            lower.append(n.lower)
            upper.append(n.upper)
        opt = self.bound_opts['optimize_bound_args']
        use_beta = bool(opt.get('enable_beta_crown', False))
        beta = self._beta_args(Bd, use_beta)
        if not optimize:
            alpha, pos = self._alpha_args(Bd, reuse_alpha)
            lb, lA = self.plan.crown_pass(C, x_L, x_U, lower, upper, alpha, pos, beta, want_lA=True)
            self.last_n_iter = 1
        else:
            alpha, pos = self._alpha_args(Bd, True)
            if alpha is None:
                raise RuntimeError("method='crown-optimized' needs alpha: call init_alpha / set the slopes first")
            rhs = decision_thresh if decision_thresh is not None else _threshold_of(opt.get('stop_criterion_func'))
            if rhs is not None:
                rhs = rhs.detach().to(dev, torch.float32).reshape(Bd, -1).contiguous()
            # the reference REPLACES alpha tensors by the best snapshots (AL/optimized_bounds.py:586):
            # work on fresh copies so that tensors the caller still holds are not modified
            for act, a in zip(self.perturbed_optimizable_activations, alpha):
                if a is not None:
                    act.alpha[self.final_name] = a.clone()
            alpha, pos = self._alpha_args(Bd, True)
            if beta is not None:
                for pre in self.layers_requiring_bounds:
                    sb = pre.sparse_betas[0] if isinstance(pre.sparse_betas, list) and pre.sparse_betas else None
                    if sb is not None:
                        sb.val = sb.val.clone()
                beta = self._beta_args(Bd, use_beta)
            lb, lA, n_iter = self.plan.optimize(
                C, x_L, x_U, lower, upper, alpha, pos, beta, rhs,
                iteration=int(opt.get('iteration', 20)), lr_alpha=float(opt.get('lr_alpha', 0.5)),
                lr_beta=float(opt.get('lr_beta', 0.05)), lr_decay=float(opt.get('lr_decay', 0.98)),
                early_stop_patience=int(opt.get('early_stop_patience', 10)),
                start_save_best=float(opt.get('start_save_best', 0.5)), enable_beta=use_beta,
                # the data-dependent exits (all verified / patience) are taken on the DEVICE (a `done` flag every kernel
                # checks): same results as breaking out of the loop, without one stream synchronisation per iteration;
                # bound_opts['optimize_bound_args']['early_stop'] = True asks for the synchronising form (exact n_iter)
                early_stop=bool(opt.get('early_stop', False)), want_lA=True)
            self.last_n_iter = n_iter
        for act, a in zip(self.perturbed_optimizable_activations, lA):
            act.lA = a                      # [S,Bd,*shape], the LAST executed pass (OP/relu.py:244)
        return lb, None
