"""Root phase and input-split regime against vectors recorded from the unmodified reference
(oracle/gen_root_golden.py): `compute_bounds(method='backward')` without interm_bounds bounds every layer exactly as
the reference does (interval arithmetic, then CROWN for the neurons it leaves unstable), and the ACAS Xu 1_1 / property 1
run of BASELINE.json configs[0] through `NetworkAbstractor.initialize` + `forward` (= `_forward_input`)."""
import os

import pytest
import torch

from fixtures import GOLDEN, load_fixture

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _close(a, b, tol=1e-5):
    b = b.to(a.device)
    return torch.allclose(a, b, rtol=tol, atol=tol * max(1.0, float(b.abs().max())))


@pytest.mark.parametrize('name', ['fc_small', 'mnist_fc', 'conv_small', 'resnet_bn_small'])
def test_full_crown_matches_reference(name):
    from neuralsat_b200.bounded_module import BoundedModule, BoundedTensor, PerturbationLpNorm
    ent = torch.load(os.path.join(GOLDEN, f'root_{name}.pt'), weights_only=False)
    fx, model, nodes = load_fixture(name)
    net = BoundedModule(model, torch.zeros(1, *nodes[0]['shape']), device=DEV)
    x_L, x_U = ent['x_L'].to(DEV), ent['x_U'].to(DEV)
    x = BoundedTensor(x_L, PerturbationLpNorm(x_L=x_L, x_U=x_U))
    lb, _ = net.compute_bounds(x=(x,), C=ent['C'].to(DEV), method='backward')
    assert _close(lb, ent['out_lb']), (lb.cpu() - ent['out_lb']).abs().max()
    for k, n in enumerate(net.split_nodes):
        assert _close(n.lower, ent['lower'][k]), (k, (n.lower.cpu() - ent['lower'][k]).abs().max())
        assert _close(n.upper, ent['upper'][k]), (k, (n.upper.cpu() - ent['upper'][k]).abs().max())
        # identical stability pattern: what the split decisions and the relaxations depend on
        ref_unstable = (ent['lower'][k] < 0) & (ent['upper'][k] > 0)
        got_unstable = ((n.lower < 0) & (n.upper > 0)).cpu()
        assert (ref_unstable != got_unstable).float().mean() < 1e-3
    for k, act in enumerate(net.perturbed_optimizable_activations):
        assert _close(act.lA, ent['lA'][k])


def test_acasxu_property_1_input_split():
    """BASELINE.json configs[0]: ONNX + VNNLIB front-ends, root bound, six generations of input bisections."""
    from neuralsat_b200.abstractor import AbstractResults, NetworkAbstractor
    from neuralsat_b200.frontend import onnx_reader, vnnlib
    d = os.path.join(GOLDEN, 'frontend')
    rec = torch.load(os.path.join(GOLDEN, 'root_acasxu.pt'), weights_only=False)
    model, in_shape, out_shape, _ = onnx_reader.parse_onnx(os.path.join(d, 'ACASXU_run2a_1_1_batch_2000.onnx'))
    obj = vnnlib.objectives(vnnlib.read_vnnlib(os.path.join(d, 'prop_1.vnnlib')))
    ab = NetworkAbstractor(model, in_shape, 'backward', input_split=True, device=DEV)
    ab.setup(obj)
    root = ab.initialize(obj)
    assert _close(root.output_lbs, rec['root_lb']), (root.output_lbs.cpu(), rec['root_lb'])
    for st in rec['steps']:
        cur = AbstractResults(objective_ids=st['ids'], input_lowers=st['in_lower'].to(DEV), input_uppers=st['in_upper'].to(DEV),
                              cs=st['cs'].to(DEV), rhs=st['rhs'].to(DEV), slopes={})
        ret = ab.forward(st['decisions'].to(DEV), cur)
        assert torch.equal(ret.input_lowers.cpu(), st['out_lower']) and torch.equal(ret.input_uppers.cpu(), st['out_upper'])
        assert _close(ret.output_lbs, st['out_lb']), (ret.output_lbs.cpu() - st['out_lb']).abs().max()
        # identical pruning decisions (NS/heuristic/domains_list.py:240-262)
        assert torch.equal((ret.output_lbs > ret.rhs).cpu(), st['out_lb'] > st['rhs'].repeat(2, 1))
        for k, n in enumerate(ab.net.split_nodes):
            assert _close(n.lower, st['lower'][k]) and _close(n.upper, st['upper'][k]), k


def test_initialize_then_hidden_split_step():
    """A verification started by the facade alone: root alpha-CROWN (output-node slopes over CROWN intermediate bounds),
    then one hidden-split BaB step from that root.  The root must be sound w.r.t. sampled outputs and at least as tight
    as plain CROWN; the children (20 iterations from the parent's fp16-rounded slopes, NS/abstractor/utils.py:51-59)
    must stay within optimiser noise of the parent's 50-iteration bound or improve on it."""
    from neuralsat_b200.abstractor import NetworkAbstractor
    from types import SimpleNamespace
    fx, model, nodes = load_fixture('fc_small')
    g = torch.Generator().manual_seed(0)
    n_box = 6
    in_shape = tuple(nodes[0]['shape'])
    x0 = torch.rand(n_box, *in_shape, generator=g)
    x_L, x_U = (x0 - 0.08).clamp(min=0), (x0 + 0.08).clamp(max=1)
    n_out = int(nodes[-1]['shape'][0])
    C = torch.randn(n_box, 1, n_out, generator=g)
    with torch.no_grad():
        y = model(x0)
    rhs = torch.einsum('bsn,bn->bs', C, y) - 0.05           # undecided at the root, provable after a few splits
    obj = SimpleNamespace(lower_bounds=x_L.flatten(1), upper_bounds=x_U.flatten(1), cs=C, rhs=rhs, ids=torch.arange(n_box) + 3)
    ab = NetworkAbstractor(model, (1, *in_shape), 'crown-optimized', input_split=False, device=DEV)
    ab.setup(obj)
    root = ab.initialize(obj)
    lb_root = root.output_lbs.to(DEV)
    # soundness against samples
    with torch.no_grad():
        t = torch.rand(64, *x_L.shape, generator=g)
        xs = x_L + t * (x_U - x_L)
        ys = model(xs.view(-1, *in_shape)).view(64, n_box, n_out)
        worst = torch.einsum('bsn,kbn->kbs', C, ys).min(0).values
    assert (lb_root.cpu() <= worst + 1e-4).all()
    if root.lower_bounds is None:
        return                                                  # everything verified at the root
    from neuralsat_b200.bounded_module import BoundedTensor, PerturbationLpNorm
    x = BoundedTensor(x_L.to(DEV), PerturbationLpNorm(x_L=x_L.to(DEV), x_U=x_U.to(DEV)))
    lb_crown, _ = ab.net.compute_bounds(x=(x,), C=C.to(DEV), method='backward',
                                        interm_bounds={k: [root.lower_bounds[k].to(DEV), root.upper_bounds[k].to(DEV)]
                                                       for k in root.lower_bounds if k != ab.net.final_name})
    assert (lb_root >= lb_crown - 1e-5).all()
    # one hidden-split step on the first unstable neuron of every domain
    name = ab.net.split_nodes[-1].name
    l, u = root.lower_bounds[name], root.upper_bounds[name]
    decisions = []
    for b in range(n_box):
        un = ((l[b] < 0) & (u[b] > 0)).flatten().nonzero().flatten()
        decisions.append([name, int(un[0]) if un.numel() else 0, 0.0])
    params = root._replace(betas=[None] * n_box, histories=[{n.name: ([], [], []) for n in ab.net.split_nodes} for _ in range(n_box)])
    ret = ab.forward(decisions, params)
    child = ret.output_lbs.view(2, n_box, -1)
    assert (child.min(0).values >= root.output_lbs.cpu() - 1e-2).all()
    assert (child.max(0).values > root.output_lbs.cpu() + 1e-4).float().mean() > 0.5       # splitting helps
