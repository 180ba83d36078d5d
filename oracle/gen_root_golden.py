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


if __name__ == '__main__':
    rb.bootstrap()
    from setting import Settings
    Settings.use_restart = False
    for name, n_box, eps in [('fc_small', 5, 0.05), ('mnist_fc', 3, 0.02), ('conv_small', 4, 0.05), ('resnet_bn_small', 3, 0.03)]:
        root_of_model(name, n_box, eps, seed=1)
    acasxu()
