"""Golden vectors for the root phase and the input-split regime, from the UNMODIFIED reference
on CPU (build container only):

    python oracle/gen_root_golden.py        # writes tests/golden/root_*.pt

* root_<model>.pt : `BoundedModule.compute_bounds(method='backward')` WITHOUT interm_bounds on a batch of boxes: the
  output bound and the bounds the reference leaves on every split node (compute_intermediate_bounds,
  AL/bound_general.py:782-903), plus `NetworkAbstractor.initialize` with method 'backward'.
* root_acasxu.pt : BASELINE.json configs[0] - the ACAS Xu 1_1 network (read by neuralsat_b200.frontend.onnx_reader from
  tests/golden/frontend) with property 1: `initialize`, then `NetworkAbstractor.forward` (= `_forward_input`,
  abstractor/abstractor.py:348-399) on widest-dimension decisions for a few generations of boxes.

TEST INFRASTRUCTURE ONLY; needs /root/reference."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from oracle import ref_bootstrap as rb  # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden')


def split_bounds(net):
    net.get_split_nodes(input_split=False)          # the split-node list is only complete after a bound computation
    return ([n.lower.detach().clone() for n in net.split_nodes], [n.upper.detach().clone() for n in net.split_nodes])


def root_of_model(name, n_box, eps, seed):
    from abstractor.abstractor import NetworkAbstractor
    from onnx2pytorch.convert.model import ConvertModel
    from fixtures import load_fixture
    fx, model, nodes = load_fixture(name)
    in_shape = tuple(nodes[0]['shape'])
    g = torch.Generator().manual_seed(seed)
    x0 = torch.rand(n_box, *in_shape, generator=g)
    r = eps * (0.5 + torch.rand(n_box, *[1] * len(in_shape), generator=g))
    x_L, x_U = (x0 - r).clamp(min=0), (x0 + r).clamp(max=1)
    n_out = int(nodes[-1]['shape'][0])
    C = torch.randn(n_box, 2, n_out, generator=g)
    rhs = torch.zeros(n_box, 2)
    obj = rb.Objective(x_L.flatten(1), x_U.flatten(1), C, rhs, torch.arange(n_box) + 3)
    ab = NetworkAbstractor(ConvertModel(model).eval(), (1, *in_shape), 'backward', input_split=False, device='cpu')
    ab.setup(obj)
    x = ab.new_input(x_L=x_L, x_U=x_U)
    with torch.no_grad():
        lb, _ = ab.net.compute_bounds(x=(x,), C=C, method='backward')
    lo, up = split_bounds(ab.net)
    ent = {'model': name, 'x_L': x_L, 'x_U': x_U, 'C': C, 'rhs': rhs, 'out_lb': lb.detach().clone(), 'lower': lo, 'upper': up,
           'lA': [m.lA.detach().clone() for m in ab.net.perturbed_optimizable_activations]}
    torch.save(ent, os.path.join(OUT, f'root_{name}.pt'))
    print(name, 'lb', lb.flatten()[:4].tolist(), 'unstable per layer',
          [int(((l < 0) & (u > 0)).sum()) for l, u in zip(lo, up)])


def acasxu():
    from abstractor.abstractor import NetworkAbstractor
    from onnx2pytorch.convert.model import ConvertModel
    from neuralsat_b200.frontend import onnx_reader, vnnlib
    d = os.path.join(OUT, 'frontend')
    model, in_shape, out_shape, _ = onnx_reader.parse_onnx(os.path.join(d, 'ACASXU_run2a_1_1_batch_2000.onnx'))
    obj = vnnlib.objectives(vnnlib.read_vnnlib(os.path.join(d, 'prop_1.vnnlib')))
    ref_obj = rb.Objective(obj.lower_bounds, obj.upper_bounds, obj.cs, obj.rhs, obj.ids)
    ab = NetworkAbstractor(ConvertModel(model).eval(), in_shape, 'backward', input_split=True, device='cpu')
    ab.setup(ref_obj)
    root = ab.initialize(ref_obj)
    rec = {'root_lb': root.output_lbs.detach().clone(), 'steps': []}
    cur = root
    for it in range(6):
        lo, up = cur.input_lowers.flatten(1), cur.input_uppers.flatten(1)
        dec = (up - lo).argmax(dim=1, keepdim=True)                  # widest dimension (decision_heuristics.py:255-261)
        ret = ab.forward(dec, cur)
        lows, ups = split_bounds(ab.net)
        rec['steps'].append({'in_lower': cur.input_lowers.clone(), 'in_upper': cur.input_uppers.clone(), 'cs': cur.cs.clone(),
                             'rhs': cur.rhs.clone(), 'ids': cur.objective_ids.clone(), 'decisions': dec.clone(),
                             'out_lb': ret.output_lbs.detach().clone(), 'out_lower': ret.input_lowers.clone(),
                             'out_upper': ret.input_uppers.clone(), 'lower': lows, 'upper': ups})
        keep = (ret.output_lbs <= ret.rhs).all(1)                    # domains that stay undecided (domains_list.py:240-262)
        print('acas step', it, 'children', len(ret.output_lbs), 'undecided', int(keep.sum()), 'min lb', float(ret.output_lbs.min()))
        if keep.sum() == 0:
            break
        from abstractor.abstractor import AbstractResults
        cur = AbstractResults(objective_ids=ret.objective_ids[keep], output_lbs=ret.output_lbs[keep],
                              input_lowers=ret.input_lowers[keep], input_uppers=ret.input_uppers[keep],
                              slopes=ret.slopes, lAs=None, cs=ret.cs[keep], rhs=ret.rhs[keep])
    torch.save(rec, os.path.join(OUT, 'root_acasxu.pt'))


def bab_run(name, batch, topk, eps, seed=0, max_iters=80):
    """The reference's hidden-split BaB loop run to the end (Verifier._parallel_dpll, NS/verifier/verifier.py:350-431,
    without attack / restart): root result, per-iteration decisions and queue lengths, final verdict."""
    import copy
    import random
    from abstractor.abstractor import NetworkAbstractor
    from abstractor.utils import new_slopes
    from heuristic.decision_heuristics import DecisionHeuristic
    from heuristic.domains_list import DomainsList
    from onnx2pytorch.convert.model import ConvertModel
    from models import build_model
    from oracle.gen_golden import _canon_names, _canon_results, flat_index
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
    cs = []
    for j in [j for j in range(n_out) if j != label]:
        c = torch.zeros(1, n_out)
        c[0, label], c[0, j] = 1., -1.
        cs.append(c)
    cs = torch.stack(cs)
    N = len(cs)
    obj = rb.Objective((x0 - eps).clamp(min=0).repeat(N, 1), (x0 + eps).clamp(max=1).repeat(N, 1), cs, torch.zeros(N, 1),
                       torch.arange(N) + 3)
    ab = NetworkAbstractor(ConvertModel(model).eval(), (1, *in_shape), 'crown-optimized', input_split=False, device='cpu')
    ab.setup(obj)
    ret = ab.initialize(obj)
    nm = _canon_names(ab.net)
    root = _canon_results(ret._replace(slopes=new_slopes(ret.slopes, ab.net.final_name), histories=None), nm)
    dl = DomainsList(net=ab.net, objective_ids=ret.objective_ids, output_lbs=ret.output_lbs, input_lowers=ret.input_lowers,
                     input_uppers=ret.input_uppers, lower_bounds=ret.lower_bounds, upper_bounds=ret.upper_bounds, lAs=ret.lAs,
                     slopes=new_slopes(ret.slopes, ab.net.final_name), histories=copy.deepcopy(ret.histories), cs=ret.cs,
                     rhs=ret.rhs, input_split=False, preconditions={})
    decision = DecisionHeuristic(input_split=False, decision_topk=topk, decision_method='smart')
    its = []
    while len(dl) > 0 and len(its) < max_iters:
        pick = dl.pick_out(batch, 'cpu')
        dec = decision(ab, pick)
        out = ab.forward(dec, pick)
        dl.add(out, dec)
        its.append({'picked': len(dec), 'remaining': len(dl), 'decisions': [(nm[d[0]], int(d[1])) for d in dec],
                    'out_lb': out.output_lbs.detach().clone()})
    verdict = 'unsat' if len(dl) == 0 else 'unknown'
    print(f'[bab {name}] {len(its)} iterations, visited {dl.visited}, verdict {verdict}, queue {[i["remaining"] for i in its]}')
    fixture = {'model': name, 'in_shape': tuple(in_shape), 'state_dict': {k: v.clone() for k, v in model.state_dict().items()},
               'root': root, 'iterations': its, 'verdict': verdict, 'visited': dl.visited, 'batch': batch, 'topk': topk,
               'alpha_index': [flat_index(getattr(m, 'alpha_indices', None), tuple(m.inputs[0].output_shape[1:]))
                               for m in ab.net.perturbed_optimizable_activations]}
    torch.save(fixture, os.path.join(OUT, f'bab_{name}.pt'))


if __name__ == '__main__':
    rb.bootstrap()
    from setting import Settings
    Settings.use_restart = False
    for name, n_box, eps in [('fc_small', 5, 0.05), ('mnist_fc', 3, 0.02), ('conv_small', 4, 0.05), ('resnet_bn_small', 3, 0.03)]:
        root_of_model(name, n_box, eps, seed=1)
    acasxu()
    bab_run('fc_small', batch=8, topk=2, eps=0.18)
    bab_run('conv_small', batch=6, topk=2, eps=0.2)
