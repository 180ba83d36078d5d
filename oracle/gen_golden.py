"""Generate golden fixtures by running the UNMODIFIED reference on CPU (build container only).

    python oracle/gen_golden.py            # writes tests/golden/*.pt

For each small network below the reference's own BaB machinery is driven exactly as
`Verifier._parallel_dpll` does (NS/verifier/verifier.py:350-431):
    NetworkAbstractor.initialize -> DomainsList -> [pick_out -> DecisionHeuristic -> abstractor.forward -> add]*
and every `BoundedModule.compute_bounds` call issued from `_forward_hidden`
(NS/abstractor/abstractor.py:280,304) is recorded: inputs (C, x_L, x_U, interm_bounds, alpha,
alpha_indices, beta loc/sign/val/bias, rhs, options) and outputs (lb, relu.lA, alpha/beta after).
Tensors are keyed by ORDER (k-th activation / k-th split node), not by the reference's tracer names.

TEST INFRASTRUCTURE ONLY; needs /root/reference (absent on the GPU box), so nothing in the
gpu tests / smoke / bench imports this file.
"""
import copy
import os
import random
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_bootstrap as rb  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


# ------------------------------------------------------------------------------------------
# model zoo (shared with tests through tests/models.py — keep definitions there)
# ------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from models import build_model, MODEL_SPECS  # noqa: E402


def flat_index(alpha_indices, shape):
    """tuple of per-dim index tensors (OP/relu.py:330-332) -> flattened neuron ids."""
    if alpha_indices is None:
        return None
    if len(alpha_indices) == 1:
        return alpha_indices[0].clone().long()
    strides = []
    s = 1
    for d in reversed(shape):
        strides.insert(0, s)
        s *= d
    idx = torch.zeros_like(alpha_indices[0])
    for i, st in zip(alpha_indices, strides):
        idx = idx + i * st
    return idx.long()


def snapshot_inputs(net, kw):
    x = kw['x'][0]
    final = net.final_name
    ent = {
        'method': kw.get('method'),
        'reuse_alpha': bool(kw.get('reuse_alpha', False)),
        'C': kw['C'].detach().clone(),
        'x_L': x.ptb.x_L.detach().clone(),
        'x_U': x.ptb.x_U.detach().clone(),
        'enable_beta': bool(net.bound_opts['optimize_bound_args']['enable_beta_crown']),
        'iteration': int(net.bound_opts['optimize_bound_args']['iteration']),
        'lr_alpha': float(net.bound_opts['optimize_bound_args']['lr_alpha']),
        'lr_beta': float(net.bound_opts['optimize_bound_args']['lr_beta']),
        'lr_decay': float(net.bound_opts['optimize_bound_args']['lr_decay']),
    }
    if kw.get('decision_thresh') is not None:
        ent['rhs'] = kw['decision_thresh'].detach().clone()
    ib = kw['interm_bounds']
    ent['lower'] = [ib[n.name][0].detach().clone() for n in net.split_nodes]
    ent['upper'] = [ib[n.name][1].detach().clone() for n in net.split_nodes]
    ent['alpha'] = []
    ent['alpha_index'] = []
    for m in net.perturbed_optimizable_activations:
        ent['alpha'].append(m.alpha[final].detach().clone())
        ai = getattr(m, 'alpha_indices', None)
        ent['alpha_index'].append(flat_index(ai, tuple(m.inputs[0].output_shape[1:])))
    if ent['enable_beta']:
        ent['beta'] = []
        for n in net.split_nodes:
            sb = n.sparse_betas[0]
            ent['beta'].append({
                'val': sb.val.detach().clone(), 'loc': sb.loc.detach().clone(),
                'sign': sb.sign.detach().clone(),
                'bias': None if sb.bias is None else sb.bias.detach().clone()})
    return ent


def snapshot_outputs(net, ret, ent):
    final = net.final_name
    ent['out_lb'] = ret[0].detach().clone()
    ent['out_lA'] = [m.lA.detach().clone() for m in net.perturbed_optimizable_activations]
    ent['out_alpha'] = [m.alpha[final].detach().clone() for m in net.perturbed_optimizable_activations]
    if ent['enable_beta']:
        ent['out_beta_val'] = [n.sparse_betas[0].val.detach().clone() for n in net.split_nodes]


def run_reference(name, batch, n_iters, topk, eps, seed=0, n_spec=1, keep=None):
    from abstractor.abstractor import NetworkAbstractor
    from heuristic.domains_list import DomainsList
    from heuristic.decision_heuristics import DecisionHeuristic
    from abstractor.utils import new_slopes
    from onnx2pytorch.convert.model import ConvertModel
    from setting import Settings
    Settings.use_restart = False

    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    model, in_shape = build_model(name)
    model.eval()
    n_in = int(np.prod(in_shape))
    x0 = torch.rand(1, n_in)
    with torch.no_grad():
        y = model(x0.view(1, *in_shape))
    n_out = y.shape[1]
    label = int(y.argmax())
    others = [j for j in range(n_out) if j != label]
    # objectives: one per adversarial class (rows e_y - e_j); optionally n_spec rows per objective
    xl = (x0 - eps).clamp(min=0)
    xu = (x0 + eps).clamp(max=1)
    cs = []
    if n_spec == 1:
        for j in others:
            c = torch.zeros(1, n_out)
            c[0, label] = 1.
            c[0, j] = -1.
            cs.append(c)
    else:
        for g in range(0, len(others) - n_spec + 1, n_spec):
            c = torch.zeros(n_spec, n_out)
            for r, j in enumerate(others[g:g + n_spec]):
                c[r, label] = 1.
                c[r, j] = -1.
            cs.append(c)
    cs = torch.stack(cs)
    N = len(cs)
    obj = rb.Objective(xl.repeat(N, 1), xu.repeat(N, 1), cs, torch.zeros(N, cs.shape[1]))

    ab = NetworkAbstractor(ConvertModel(model).eval(), (1, *in_shape), 'crown-optimized',
                           input_split=False, device='cpu')
    ab.setup(obj)
    ret = ab.initialize(obj)
    if ret.slopes is None:
        raise RuntimeError(f'{name}: verified at the root; raise eps')
    dl = DomainsList(net=ab.net, objective_ids=ret.objective_ids, output_lbs=ret.output_lbs,
                     input_lowers=ret.input_lowers, input_uppers=ret.input_uppers,
                     lower_bounds=ret.lower_bounds, upper_bounds=ret.upper_bounds, lAs=ret.lAs,
                     slopes=new_slopes(ret.slopes, ab.net.final_name),
                     histories=copy.deepcopy(ret.histories), cs=ret.cs, rhs=ret.rhs,
                     input_split=False, preconditions={})
    print(f'[{name}] root lbs {ret.output_lbs.flatten().tolist()} -> {len(dl)} domains')
    decision = DecisionHeuristic(input_split=False, decision_topk=topk, decision_method='smart')

    records = []
    orig = ab.net.compute_bounds

    depth = [0]

    def rec(*args, **kw):
        # record only the outermost call (the optimiser re-enters compute_bounds per iteration)
        if depth[0] > 0 or args or kw.get('interm_bounds') is None or 'x' not in kw:
            return orig(*args, **kw)
        depth[0] += 1
        try:
            ent = snapshot_inputs(ab.net, kw)
            out = orig(*args, **kw)
            snapshot_outputs(ab.net, out, ent)
            records.append(ent)
        finally:
            depth[0] -= 1
        return out

    ab.net.compute_bounds = rec
    decisions_log = []
    for it in range(n_iters):
        if len(dl) == 0:
            break
        pick = dl.pick_out(batch, 'cpu')
        dec = decision(ab, pick)
        decisions_log.append([(ab.net.split_nodes.index(ab.net[d[0]]), int(d[1]), float(d[2])) for d in dec])
        out = ab.forward(dec, pick)
        dl.add(out, dec)
        print(f'[{name}] iter {it}: picked {len(dec)}, remaining {len(dl)}, records {len(records)}')
    ab.net.compute_bounds = orig

    f1 = [r for r in records if r['method'] == 'backward']
    f2 = [r for r in records if r['method'] == 'crown-optimized']
    if keep is not None:
        f1 = f1[-keep[0]:] if keep[0] else []
        f2 = f2[-keep[1]:] if keep[1] else []
    fixture = {
        'model': name, 'in_shape': tuple(in_shape), 'seed': seed, 'eps': eps,
        'state_dict': {k: v.clone() for k, v in model.state_dict().items()},
        'f1': f1, 'f2': f2, 'decisions': decisions_log,
        'reference': 'dynaroars/neuralsat @916eb56 (neuralsat-pt201), torch ' + torch.__version__,
    }
    path = os.path.join(OUT, f'{name}.pt')
    torch.save(fixture, path)
    print(f'[{name}] wrote {path}: {len(f1)} F1 + {len(f2)} F2 records, {os.path.getsize(path)/1e6:.2f} MB')


def _canon_names(net):
    """reference tracer names -> order-based keys shared with the facade ('pre{k}', 'act{k}', 'final')."""
    m = {net.final_name: 'final'}
    for k, n in enumerate(net.split_nodes):
        m[n.name] = f'pre{k}'
    for k, a in enumerate(net.perturbed_optimizable_activations):
        m[a.name] = f'act{k}'
    return m


def _canon_results(r, nm):
    """AbstractResults -> plain dict with canonical keys (tensors cloned to CPU)."""
    def t(v):
        return None if v is None else v.detach().to('cpu').clone()
    d = {}
    for f in ('objective_ids', 'output_lbs', 'input_lowers', 'input_uppers', 'cs', 'rhs'):
        d[f] = t(getattr(r, f))
    for f in ('lAs', 'lower_bounds', 'upper_bounds'):
        v = getattr(r, f)
        d[f] = None if v is None else {nm[k]: t(x) for k, x in v.items()}
    d['slopes'] = None if r.slopes is None else {nm[k]: {nm[kk]: t(x) for kk, x in v.items()} for k, v in r.slopes.items()}
    d['betas'] = None if r.betas is None else [None if b is None else {nm[k]: t(x) for k, x in b.items()} for b in r.betas]
    d['histories'] = None if r.histories is None else [
        {nm[k]: tuple(torch.as_tensor(x).clone() for x in v) for k, v in h.items()} for h in r.histories]
    return d


def run_abstractor_records(name, batch, n_iters, topk, eps, seed=0, keep_abs=2, **_):
    """Records NetworkAbstractor.forward(decisions, domain_params) -> AbstractResults of the last
    `keep_abs` BaB iterations of the reference (NS/abstractor/abstractor.py:403-406, :244-344):
    the boundary the facade `neuralsat_b200.abstractor.NetworkAbstractor` mirrors."""
    from abstractor.abstractor import NetworkAbstractor
    from heuristic.domains_list import DomainsList
    from heuristic.decision_heuristics import DecisionHeuristic
    from abstractor.utils import new_slopes
    from onnx2pytorch.convert.model import ConvertModel
    from setting import Settings
    Settings.use_restart = False
    torch.manual_seed(seed)
    np.random.seed(seed)
    random.seed(seed)
    model, in_shape = build_model(name)
    model.eval()
    n_in = int(np.prod(in_shape))
    x0 = torch.rand(1, n_in)
    with torch.no_grad():
        y = model(x0.view(1, *in_shape))
    n_out = y.shape[1]
    label = int(y.argmax())
    others = [j for j in range(n_out) if j != label]
    xl = (x0 - eps).clamp(min=0)
    xu = (x0 + eps).clamp(max=1)
    cs = []
    for j in others:
        c = torch.zeros(1, n_out)
        c[0, label] = 1.
        c[0, j] = -1.
        cs.append(c)
    cs = torch.stack(cs)
    N = len(cs)
    obj = rb.Objective(xl.repeat(N, 1), xu.repeat(N, 1), cs, torch.zeros(N, cs.shape[1]))
    ab = NetworkAbstractor(ConvertModel(model).eval(), (1, *in_shape), 'crown-optimized', input_split=False, device='cpu')
    ab.setup(obj)
    ret = ab.initialize(obj)
    dl = DomainsList(net=ab.net, objective_ids=ret.objective_ids, output_lbs=ret.output_lbs,
                     input_lowers=ret.input_lowers, input_uppers=ret.input_uppers,
                     lower_bounds=ret.lower_bounds, upper_bounds=ret.upper_bounds, lAs=ret.lAs,
                     slopes=new_slopes(ret.slopes, ab.net.final_name),
                     histories=copy.deepcopy(ret.histories), cs=ret.cs, rhs=ret.rhs,
                     input_split=False, preconditions={})
    decision = DecisionHeuristic(input_split=False, decision_topk=topk, decision_method='smart')
    nm = _canon_names(ab.net)
    recs = []
    for it in range(n_iters):
        if len(dl) == 0:
            break
        pick = dl.pick_out(batch, 'cpu')
        dec = decision(ab, pick)
        rec = {'params': _canon_results(pick, nm), 'decisions': [(nm[d[0]], int(d[1]), float(d[2])) for d in dec]}
        out = ab.forward(dec, pick)
        rec['out'] = _canon_results(out, nm)
        rec['alpha_index'] = [flat_index(getattr(m, 'alpha_indices', None), tuple(m.inputs[0].output_shape[1:]))
                              for m in ab.net.perturbed_optimizable_activations]
        recs.append(rec)
        dl.add(out, dec)
        print(f'[{name}/abs] iter {it}: picked {len(dec)}, remaining {len(dl)}')
    fixture = {'model': name, 'in_shape': tuple(in_shape), 'seed': seed, 'eps': eps,
               'state_dict': {k: v.clone() for k, v in model.state_dict().items()},
               'records': recs[-keep_abs:],
               'reference': 'dynaroars/neuralsat @916eb56 (neuralsat-pt201), torch ' + torch.__version__}
    path = os.path.join(OUT, f'{name}_abs.pt')
    torch.save(fixture, path)
    print(f'[{name}/abs] wrote {path}: {len(fixture["records"])} records, {os.path.getsize(path)/1e6:.2f} MB')


def run_toy_known_answers():
    """Known-answer vectors of NS/example/test_model.py:80-108 (fixed-weight ReLUNet), re-derived
    by running the reference here; compared against SURVEY.md section 8c in the test."""
    from auto_LiRPA import BoundedModule, BoundedTensor
    from auto_LiRPA.perturbations import PerturbationLpNorm
    model, in_shape = build_model('toy_fixed')
    x_L = torch.tensor([[-1., -2.]])
    x_U = torch.tensor([[1., 2.]])
    net = BoundedModule(model, torch.zeros(1, 2), bound_opts={'relu': 'adaptive', 'conv_mode': 'matrix'})
    x = BoundedTensor((x_L + x_U) / 2, PerturbationLpNorm(x_L=x_L, x_U=x_U))
    lb, ub = net.compute_bounds(x=(x,), method='backward', bound_upper=True)
    C = torch.tensor([[[1., -1., 0.]]])
    lbc, _ = net.compute_bounds(x=(x,), C=C, method='backward', bound_upper=False)
    pre = [(n.lower.detach().clone(), n.upper.detach().clone()) for n in net.get_split_nodes()[0]]
    fixture = {'lb': lb.detach(), 'ub': ub.detach(), 'lb_C': lbc.detach(), 'pre': pre,
               'x_L': x_L, 'x_U': x_U, 'C': C}
    torch.save(fixture, os.path.join(OUT, 'toy_fixed.pt'))
    print('[toy_fixed]', lb.tolist(), ub.tolist(), lbc.tolist(), [(l.tolist(), u.tolist()) for l, u in pre])


if __name__ == '__main__':
    rb.bootstrap()
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ['toy_fixed'] + list(MODEL_SPECS)
    for name in which:
        if name == 'toy_fixed':
            run_toy_known_answers()
        elif name.endswith('_abs'):
            run_abstractor_records(name[:-4], **MODEL_SPECS[name[:-4]])
        else:
            run_reference(name, **MODEL_SPECS[name])
