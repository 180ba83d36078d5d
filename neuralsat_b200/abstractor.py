"""Host-side mirror of `NetworkAbstractor` (NS/abstractor/abstractor.py:22-423) for the hidden-split
BaB loop: same method names, argument meaning and result layout (`AbstractResults`), with
`self.net` being `neuralsat_b200.BoundedModule` (CUDA library underneath).

Mirrored (paths under /root/reference/neuralsat-pt201):
  NetworkAbstractor.forward / _forward_hidden ...... abstractor/abstractor.py:403-406, :244-344
  new_input, get_slope, set_slope, get_lAs ......... abstractor/utils.py:24-27, :51-74, :98-106
  get_hidden_bounds, get_beta, reset_beta, set_beta  abstractor/utils.py:78-94, :110-131, :182-209
  update_histories, hidden_split_idx ............... abstractor/utils.py:159-178, :214-250
  get_branching_opt_params / get_beta_opt_params ... abstractor/params.py:21-28, :51-65
  AbstractResults .................................. util/misc/result.py:4-17

  NetworkAbstractor.initialize ...................... abstractor/abstractor.py:153-240
  NetworkAbstractor._forward_input, input_split_idx  abstractor/abstractor.py:348-399, abstractor/utils.py:255-280
  get_initialize_opt_params / get_input_opt_params .. abstractor/params.py:32-48, :68-82

Root bounds (`initialize`) and the input-split regime (`_forward_input`) bound every intermediate layer by CROWN
(`BoundedModule._intermediate_bounds`).  With method 'backward' this is the reference's computation; with
'crown-optimized' the reference additionally re-tightens the intermediate layers in each of its 50 root iterations,
here only the output node's slopes are optimised over fixed CROWN intermediate bounds (sound, looser root).
"""
from __future__ import annotations

import copy
from collections import namedtuple
from typing import Dict, List

import torch

from .bounded_module import BoundedModule, BoundedTensor, PerturbationLpNorm, stop_criterion_batch_any

AbstractResults = namedtuple(
    'AbstractResults',
    ('objective_ids', 'output_lbs', 'masks', 'lAs', 'histories', 'lower_bounds', 'upper_bounds',
     'input_lowers', 'input_uppers', 'slopes', 'betas', 'cs', 'rhs', 'sat_solvers'),
    defaults=(None,) * 14)

BACKWARD_BATCH_SIZE = 10 ** 9       # Settings.backward_batch_size; irrelevant here (no crown batching)


def get_branching_opt_params() -> dict:
    return {'crown_batch_size': BACKWARD_BATCH_SIZE,
            'optimize_bound_args': {'enable_beta_crown': False, 'fix_interm_bounds': True}}


def get_initialize_opt_params(stop_criterion_func) -> dict:
    return {'crown_batch_size': BACKWARD_BATCH_SIZE,
            'optimize_bound_args': {'enable_alpha_crown': True, 'enable_beta_crown': False, 'use_shared_alpha': False,
                                    'init_alpha': False, 'fix_interm_bounds': True,
                                    'stop_criterion_func': stop_criterion_func, 'iteration': 50, 'lr_alpha': 0.1,
                                    'lr_beta': 0.1, 'lr_decay': 0.98}}


def get_input_opt_params(stop_criterion_func) -> dict:
    return {'crown_batch_size': BACKWARD_BATCH_SIZE,
            'optimize_bound_args': {'enable_beta_crown': False, 'fix_interm_bounds': True, 'iteration': 20,
                                    'lr_alpha': 0.1, 'lr_decay': 0.98, 'stop_criterion_func': stop_criterion_func}}


def get_beta_opt_params(stop_criterion_func) -> dict:
    return {'crown_batch_size': BACKWARD_BATCH_SIZE,
            'optimize_bound_args': {'enable_alpha_crown': True, 'enable_beta_crown': True,
                                    'use_shared_alpha': False, 'fix_interm_bounds': True, 'iteration': 20,
                                    'lr_alpha': 0.1, 'lr_beta': 0.1, 'lr_decay': 0.98,
                                    'stop_criterion_func': stop_criterion_func}}


def _to_device(t: torch.Tensor, device='cpu', half=False) -> torch.Tensor:
    if half:
        t = t.half()
    return t.to(device)


def _append(t, value, dtype=torch.float32) -> torch.Tensor:
    if not isinstance(t, torch.Tensor):
        t = torch.tensor(t, dtype=dtype)
    out = torch.empty(len(t) + 1, dtype=dtype)
    out[:len(t)] = t
    out[-1] = value
    return out


class NetworkAbstractor:
    """alpha-beta-CROWN abstraction of one network on one GPU."""

    def __init__(self, pytorch_model, input_shape: tuple, method: str = 'crown-optimized',
                 input_split: bool = False, device: str = 'cuda'):
        self.pytorch_model = copy.deepcopy(pytorch_model)
        self.device = device
        self.input_shape = tuple(input_shape)
        self.input_split = input_split
        self.method = method
        self.mode = 'matrix'
        self.iteration = 0
        self.net = BoundedModule(self.pytorch_model, torch.zeros(self.input_shape), device=device)
        self.net.get_split_nodes(input_split=False)

    @property
    def split_points(self):
        if not hasattr(self, '_split_points'):
            self._split_points = [self.net.split_activations[k.name][0][0].get_split_point()
                                  for k in self.net.split_nodes]
        return self._split_points

    def setup(self, objective=None) -> None:
        return None

    def initialize(self, objective, reference_bounds=None, init_betas=None) -> AbstractResults:
        """Root bounds of a batch of objectives (abstractor/abstractor.py:153-240)."""
        objective.cs = objective.cs.to(self.device)
        objective.rhs = objective.rhs.to(self.device)
        input_lowers = objective.lower_bounds.view(-1, *self.input_shape[1:]).to(self.device)
        input_uppers = objective.upper_bounds.view(-1, *self.input_shape[1:]).to(self.device)
        stop_criterion_func = stop_criterion_batch_any(objective.rhs)
        x = self.new_input(x_L=input_lowers, x_U=input_uppers)
        self.init_reference_bounds = reference_bounds
        self.net.get_split_nodes(input_split=False)
        if self.method not in ('crown-optimized',):
            lb, _ = self.net.compute_bounds(x=(x,), C=objective.cs, method=self.method, reference_bounds=reference_bounds)
            if stop_criterion_func(lb).all().item():
                return AbstractResults(output_lbs=lb)
            return AbstractResults(objective_ids=getattr(objective, 'ids', None), output_lbs=lb, slopes=self.get_slope(),
                                   lAs=self.get_lAs(), cs=objective.cs, rhs=objective.rhs,
                                   input_lowers=input_lowers, input_uppers=input_uppers)
        self.net.set_bound_opts(get_initialize_opt_params(stop_criterion_func))
        lb, _, aux_reference_bounds = self.net.init_alpha(x=(x,), share_alphas=False, c=objective.cs, bound_upper=False)
        if stop_criterion_func(lb).all().item():
            return AbstractResults(output_lbs=lb)
        lb, _ = self.net.compute_bounds(x=(x,), C=objective.cs, method='crown-optimized',
                                        aux_reference_bounds=aux_reference_bounds, reference_bounds=reference_bounds)
        if stop_criterion_func(lb).all().item():
            return AbstractResults(output_lbs=lb)
        with torch.no_grad():
            lower_bounds, upper_bounds = self.get_hidden_bounds(lb)
        return AbstractResults(objective_ids=objective.ids, output_lbs=lower_bounds[self.net.final_name],
                               lAs=self.get_lAs(), lower_bounds=lower_bounds, upper_bounds=upper_bounds,
                               slopes=self.get_slope(),
                               histories={n.name: ([], [], []) for n in self.net.split_nodes},
                               cs=objective.cs, rhs=objective.rhs, input_lowers=input_lowers, input_uppers=input_uppers)

    def __repr__(self):
        return f'{self.__class__.__name__}({self.mode}, {self.method})'

    # ---- helpers (abstractor/utils.py) ---------------------------------------------------------
    def new_input(self, x_L: torch.Tensor, x_U: torch.Tensor) -> BoundedTensor:
        return BoundedTensor(x_L, PerturbationLpNorm(x_L=x_L, x_U=x_U)).to(self.device)

    def get_slope(self, half=True, device='cpu') -> dict:
        return {m.name: {name: _to_device(a, device=device, half=half) for name, a in m.alpha.items()}
                for m in self.net.perturbed_optimizable_activations}

    def set_slope(self, slope: dict) -> None:
        for m in self.net.perturbed_optimizable_activations:
            names = list(m.alpha.keys()) if m.alpha else list(slope.get(m.name, {}).keys())
            for node_name in names:
                if node_name in slope[m.name]:
                    a = slope[m.name][node_name]
                    if a.size(2) > 0:
                        a = a.to(self.device, torch.float32)
                        m.alpha[node_name] = a.repeat(1, 1, 2, *([1] * (a.ndim - 3))).contiguous()   # 2 * batch
                else:
                    del m.alpha[node_name]        # do not use alphas of other start nodes

    def get_hidden_bounds(self, output_lbs: torch.Tensor, device='cpu'):
        lower_bounds, upper_bounds = {}, {}
        for layer in self.net.split_nodes:
            lower_bounds[layer.name] = _to_device(layer.lower.detach(), device=device)
            upper_bounds[layer.name] = _to_device(layer.upper.detach(), device=device)
        lower_bounds[self.net.final_name] = _to_device(output_lbs.flatten(1).detach(), device=device)
        upper_bounds[self.net.final_name] = _to_device((output_lbs + torch.inf).flatten(1).detach(), device=device)
        return lower_bounds, upper_bounds

    def get_lAs(self, size=None, device='cpu') -> dict:
        lAs = {}
        for node in self.net.get_splittable_activations():
            if getattr(node, 'lA', None) is not None:
                lAs[node.name] = _to_device(node.lA.transpose(0, 1), device=device)
        return lAs

    def get_beta(self, num_splits: List[dict], device='cpu') -> list:
        """Per-domain {layer: beta[:n_splits]} (abstractor/utils.py:119-131).  One masked gather + one split per layer;
        the per-domain tensors are views of it."""
        if not num_splits:
            return []
        n = len(num_splits)
        out = [dict() for _ in range(n)]
        for k in num_splits[0]:
            vals = self.net[k].sparse_betas[0].val.detach().to(device)
            lens = [num_splits[i][k] for i in range(n)]
            keep = torch.arange(vals.shape[1]).unsqueeze(0) < torch.tensor(lens).unsqueeze(1)
            pieces = torch.split(vals[keep], lens)
            for i in range(n):
                out[i][k] = pieces[i]
        return out

    def reset_beta(self, batch: int, max_splits_per_layer: dict, betas=None, bias=False) -> None:
        for layer_name, width in max_splits_per_layer.items():
            layer = self.net[layer_name]
            if betas is not None and betas[0] is not None and layer_name in betas[0]:
                betas_ = [(betas[bi][layer_name] if betas[bi] is not None else None) for bi in range(batch)]
            else:
                betas_ = [None] * batch
            self.net.reset_beta(layer, (batch, width), betas_, bias=bias)

    def update_histories(self, histories: List[dict], decisions: List[list]) -> List[dict]:
        """Children [0, B) get (loc, +1, point) appended to the decision layer's history, children [B, 2B) (loc, -1, point)
        (abstractor/utils.py:159-178)."""
        batch = len(decisions)
        ids = torch.tensor([int(d[1]) for d in decisions], dtype=torch.long)
        pts = torch.tensor([float(d[2]) for d in decisions], dtype=torch.float32)
        signs = (torch.tensor([1.0]), torch.tensor([-1.0]))
        double = [dict(h) for _ in range(2) for h in histories]
        for i, h in enumerate(double):
            j = i % batch
            name = decisions[j][0]
            loc, sign, point = (t if isinstance(t, torch.Tensor) else torch.as_tensor(t) for t in h[name])
            h[name] = (torch.cat([loc.long(), ids[j:j + 1]]), torch.cat([sign.float(), signs[0 if i < batch else 1]]),
                       torch.cat([point.float(), pts[j:j + 1]]))
        return double

    def set_beta(self, betas: list, histories: List[dict]) -> List[dict]:
        batch = len(histories)
        splits_per_example, max_splits = [], {}
        for bi in range(batch):
            splits_per_example.append({k: len(v[0]) for k, v in histories[bi].items()})
            for k, n in splits_per_example[bi].items():
                max_splits[k] = max(max_splits.get(k, 0), n)
        self.reset_beta(betas=betas, max_splits_per_layer=max_splits, batch=batch, bias=None in self.split_points)
        for node in self.net.split_nodes:
            if node.sparse_betas is None:
                continue
            for sb in node.sparse_betas:
                sb.apply_splits(histories, node.name)
        return splits_per_example

    @torch.no_grad()
    def hidden_split_idx(self, lower_bounds: dict, upper_bounds: dict, decisions: List[list]) -> dict:
        """Child i (first half): lower[layer][i, n] = p; child i+B (second half): upper[layer][i+B, n] = p."""
        batch = len(decisions)
        rows: Dict[str, list] = {k: [] for k in lower_bounds}
        cols: Dict[str, list] = {k: [] for k in lower_bounds}
        pts: Dict[str, list] = {k: [] for k in lower_bounds}
        for i, (name, nid, point) in enumerate(decisions):
            rows[name].append(i)
            cols[name].append(int(nid))
            pts[name].append(float(point))
        out = {}
        for key in lower_bounds:
            lo = torch.cat([lower_bounds[key], lower_bounds[key]], dim=0).to(self.device)
            up = torch.cat([upper_bounds[key], upper_bounds[key]], dim=0).to(self.device)
            if rows[key]:
                r = torch.as_tensor(rows[key], device=self.device)
                c = torch.as_tensor(cols[key], device=self.device)
                p = torch.as_tensor(pts[key], device=self.device, dtype=lo.dtype)
                lo.view(2 * batch, -1)[r, c] = p
                up.view(2 * batch, -1)[r + batch, c] = p
            out[key] = [lo, up]
        return out

    @torch.no_grad()
    def input_split_idx(self, input_lowers: torch.Tensor, input_uppers: torch.Tensor, split_idx: torch.Tensor):
        """Bisect dimension split_idx[:, 0] of every box at its mid-point: first half keeps the upper part,
        second half the lower part (abstractor/utils.py:255-280)."""
        lo, up = input_lowers.flatten(1), input_uppers.flatten(1)
        rows = torch.arange(lo.shape[0], device=lo.device)
        idx = split_idx[:, 0].long().to(lo.device)
        mid = (lo[rows, idx] + up[rows, idx]) / 2
        lo_hi, up_lo = lo.clone(), up.clone()
        lo_hi[rows, idx] = mid
        up_lo[rows, idx] = mid
        new_lo = torch.cat([lo_hi, lo]).reshape(-1, *self.input_shape[1:])
        new_up = torch.cat([up, up_lo]).reshape(-1, *self.input_shape[1:])
        return new_lo, new_up

    def _forward_input(self, domain_params: AbstractResults, decisions: torch.Tensor, simplify: bool) -> AbstractResults:
        """Input-split step (abstractor/abstractor.py:348-399): every child box has all its layers re-bounded."""
        batch = len(decisions)
        assert batch > 0 and batch == len(domain_params.cs) == len(domain_params.rhs) == len(domain_params.input_lowers)
        new_input_lowers, new_input_uppers = self.input_split_idx(domain_params.input_lowers.to(self.device),
                                                                  domain_params.input_uppers.to(self.device), decisions)
        new_x = self.new_input(x_L=new_input_lowers, x_U=new_input_uppers)
        double_objective_ids = torch.cat([domain_params.objective_ids, domain_params.objective_ids], dim=0)
        double_cs = torch.cat([domain_params.cs, domain_params.cs], dim=0)
        double_rhs = torch.cat([domain_params.rhs, domain_params.rhs], dim=0)
        if domain_params.slopes is not None and len(domain_params.slopes) > 0:
            self.set_slope(domain_params.slopes)
        self.net.set_bound_opts(get_input_opt_params(stop_criterion_batch_any(double_rhs)))
        for n in self.net.split_nodes:
            n.lower = n.upper = None                  # bounds of the parents do not carry over to other boxes
        double_output_lbs, _ = self.net.compute_bounds(x=(new_x,), C=double_cs, method=self.method,
                                                       decision_thresh=double_rhs,
                                                       reference_bounds=getattr(self, 'init_reference_bounds', None))
        with torch.no_grad():
            double_slopes = self.get_slope() if domain_params.slopes is not None and len(domain_params.slopes) > 0 else {}
            double_lAs = self.get_lAs(size=len(new_input_lowers))
        return AbstractResults(objective_ids=double_objective_ids, output_lbs=double_output_lbs,
                               input_lowers=new_input_lowers, input_uppers=new_input_uppers, slopes=double_slopes,
                               lAs=double_lAs, cs=double_cs, rhs=double_rhs)

    # ---- the BaB step --------------------------------------------------------------------------
    def forward(self, decisions, domain_params: AbstractResults) -> AbstractResults:
        self.iteration += 1
        forward_func = self._forward_input if self.input_split else self._forward_hidden
        return forward_func(domain_params=domain_params, decisions=decisions, simplify=False)

    def _forward_hidden(self, domain_params: AbstractResults, decisions: list, simplify: bool) -> AbstractResults:
        """One hidden-split step (abstractor/abstractor.py:244-344): every picked domain yields two children - the
        decision neuron forced active in rows [0, B), inactive in rows [B, 2B) - which are bounded by one CROWN pass
        (`simplify`, the branching look-ahead) or by the alpha/beta optimisation."""
        n_parents = len(decisions)
        p = domain_params
        if n_parents == 0 or n_parents != len(p.cs) or n_parents != len(p.input_lowers):
            raise ValueError('one decision per picked domain is required')

        def twice(t):
            return torch.cat([t, t], dim=0)

        spec, box_lo, box_hi = twice(p.cs), twice(p.input_lowers), twice(p.input_uppers)
        child_bounds = self.hidden_split_idx(lower_bounds=p.lower_bounds, upper_bounds=p.upper_bounds, decisions=decisions)
        x = self.new_input(x_L=box_lo, x_U=box_hi)
        has_slopes = p.slopes is not None and len(p.slopes) > 0
        if has_slopes:
            self.set_slope(p.slopes)
        if simplify:
            self.net.set_bound_opts(get_branching_opt_params())
            lbs, _ = self.net.compute_bounds(x=(x,), C=spec, method='backward', reuse_alpha=True, interm_bounds=child_bounds)
            return AbstractResults(output_lbs=lbs)

        thresholds = twice(p.rhs)
        histories = self.update_histories(histories=p.histories, decisions=decisions)
        n_split = self.set_beta(betas=p.betas * 2, histories=histories)
        self.net.set_bound_opts(get_beta_opt_params(stop_criterion_batch_any(thresholds)))
        lbs, _ = self.net.compute_bounds(x=(x,), C=spec, method=self.method, decision_thresh=thresholds,
                                         interm_bounds=child_bounds)
        with torch.no_grad():
            lbs = lbs.detach().to('cpu')
            lower_bounds, upper_bounds = self.get_hidden_bounds(lbs)
            out = AbstractResults(
                objective_ids=twice(p.objective_ids), output_lbs=lower_bounds[self.net.final_name],
                input_lowers=box_lo, input_uppers=box_hi, lAs=self.get_lAs(size=len(box_lo)),
                lower_bounds=lower_bounds, upper_bounds=upper_bounds, slopes=self.get_slope() if has_slopes else {},
                betas=self.get_beta(n_split), histories=histories, cs=spec, rhs=thresholds,
                sat_solvers=p.sat_solvers * 2 if p.sat_solvers is not None else None)
        return out
