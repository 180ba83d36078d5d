"""GPU: the reference-facing facade (neuralsat_b200.BoundedModule / NetworkAbstractor) against golden
records of the unmodified reference: compute_bounds-level (tests/golden/<net>.pt) and
NetworkAbstractor.forward-level (tests/golden/<net>_abs.pt).  Tolerance: 1e-5 relative on bounds
(north star), identical verdicts; optimisation variables as in test_cuda_parity.test_f2_vs_reference."""
import os

import pytest
import torch

from fixtures import GOLDEN, load_fixture
from models import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _scale(t):
    return max(1.0, float(t.abs().max()))


def _module(name, state_dict):
    from neuralsat_b200 import BoundedModule
    model, in_shape = build_model(name)
    model.load_state_dict(state_dict)
    return BoundedModule(model.eval(), torch.zeros(1, *in_shape), device=DEV), in_shape


def _install(net, ent, with_beta):
    from neuralsat_b200.bounded_module import SparseBeta
    for k, act in enumerate(net.perturbed_optimizable_activations):
        act.alpha = {net.final_name: ent['alpha'][k].clone()}
        act.alpha_indices = ent['alpha_index'][k]
        act._alpha_pos = None
    for k, pre in enumerate(net.split_nodes):
        pre.sparse_betas = None
        if with_beta:
            b = ent['beta'][k]
            sb = SparseBeta(tuple(b['val'].shape), bias=b['bias'] is not None, device=DEV)
            sb.val, sb.loc, sb.sign = b['val'].to(DEV), b['loc'].to(DEV), b['sign'].to(DEV)
            sb.bias = None if b['bias'] is None else b['bias'].to(DEV)
            pre.sparse_betas = [sb]
    return {pre.name: [ent['lower'][k], ent['upper'][k]] for k, pre in enumerate(net.split_nodes)}


@pytest.mark.parametrize('name', ['fc_small', 'mnist_fc', 'conv_small', 'resnet_bn_small', 'fc_sigmoid', 'fc_tanh', 'fc_const'])
def test_compute_bounds_matches_reference(name):
    from neuralsat_b200 import BoundedTensor, PerturbationLpNorm
    from neuralsat_b200.bounded_module import stop_criterion_batch_any
    fx, _, _ = load_fixture(name)
    net, _ = _module(fx['model'], fx['state_dict'])
    for ent in fx['f1']:
        ib = _install(net, ent, False)
        x = BoundedTensor(ent['x_L'], PerturbationLpNorm(x_L=ent['x_L'], x_U=ent['x_U']))
        net.set_bound_opts({'optimize_bound_args': {'enable_beta_crown': False, 'fix_interm_bounds': True}})
        lb, ub = net.compute_bounds(x=(x,), C=ent['C'], method='backward', reuse_alpha=True, interm_bounds=ib)
        assert ub is None
        ref = ent['out_lb']
        assert torch.allclose(lb.cpu(), ref, rtol=1e-5, atol=1e-5 * _scale(ref))
        for k, act in enumerate(net.perturbed_optimizable_activations):
            assert torch.allclose(act.lA.cpu(), ent['out_lA'][k], rtol=1e-5, atol=1e-5 * _scale(ent['out_lA'][k]))
    for ent in fx['f2']:
        ib = _install(net, ent, ent['enable_beta'])
        held = [a.alpha[net.final_name] for a in net.perturbed_optimizable_activations]
        before = [h.clone() for h in held]
        x = BoundedTensor(ent['x_L'], PerturbationLpNorm(x_L=ent['x_L'], x_U=ent['x_U']))
        net.set_bound_opts({'optimize_bound_args': {
            'enable_alpha_crown': True, 'enable_beta_crown': ent['enable_beta'], 'iteration': ent['iteration'],
            'lr_alpha': ent['lr_alpha'], 'lr_beta': ent['lr_beta'], 'lr_decay': ent['lr_decay'],
            'stop_criterion_func': stop_criterion_batch_any(ent['rhs'])}})
        lb, _ = net.compute_bounds(x=(x,), C=ent['C'], method='crown-optimized', decision_thresh=ent['rhs'],
                                   interm_bounds=ib)
        ref = ent['out_lb']
        assert torch.allclose(lb.cpu(), ref, rtol=1e-5, atol=1e-5 * _scale(ref)), (lb.cpu() - ref).abs().max()
        assert torch.equal(lb.cpu() > ent['rhs'], ref > ent['rhs'])
        # alpha tensors are REPLACED, not updated in place (AL/optimized_bounds.py:586)
        for h, b0, act in zip(held, before, net.perturbed_optimizable_activations):
            assert torch.equal(h, b0)
            new = act.alpha[net.final_name]
            assert new.shape == h.shape and new.data_ptr() != h.data_ptr()


def test_unsupported_calls_fail_loudly():
    from neuralsat_b200 import BoundedTensor, PerturbationLpNorm
    fx, _, _ = load_fixture('fc_small')
    net, _ = _module(fx['model'], fx['state_dict'])
    ent = fx['f1'][0]
    x = BoundedTensor(ent['x_L'], PerturbationLpNorm(x_L=ent['x_L'], x_U=ent['x_U']))
    with pytest.raises(NotImplementedError):
        net.compute_bounds(x=(x,), C=ent['C'], method='forward', interm_bounds={})          # forward-mode LiRPA
    with pytest.raises(NotImplementedError):
        net.compute_bounds(x=(x,), C=ent['C'], method='backward', bound_upper=True)         # upper bounds of the output
    with pytest.raises(NotImplementedError):
        net.compute_bounds(x=(x,), C=ent['C'], method='backward', IBP=True)


@pytest.mark.parametrize('name', ['fc_small', 'conv_small', 'fc_sigmoid'])
def test_abstractor_forward_matches_reference(name):
    """NetworkAbstractor.forward(decisions, domain_params) of one BaB iteration, field by field."""
    from neuralsat_b200.abstractor import AbstractResults, NetworkAbstractor
    fx = torch.load(os.path.join(GOLDEN, f'{name}_abs.pt'), weights_only=False)
    model, in_shape = build_model(fx['model'])
    model.load_state_dict(fx['state_dict'])
    ab = NetworkAbstractor(model.eval(), (1, *in_shape), 'crown-optimized', input_split=False, device=DEV)
    net = ab.net
    to_name = {'final': net.final_name}
    to_name.update({f'pre{k}': n.name for k, n in enumerate(net.split_nodes)})
    to_name.update({f'act{k}': a.name for k, a in enumerate(net.perturbed_optimizable_activations)})

    def ren(d):
        return None if d is None else {to_name[k]: v for k, v in d.items()}

    for rec in fx['records']:
        p, out = rec['params'], rec['out']
        for k, act in enumerate(net.perturbed_optimizable_activations):
            act.alpha_indices, act._alpha_pos, act.alpha = rec['alpha_index'][k], None, {}
        params = AbstractResults(
            objective_ids=p['objective_ids'], output_lbs=p['output_lbs'], input_lowers=p['input_lowers'],
            input_uppers=p['input_uppers'], lAs=ren(p['lAs']), lower_bounds=ren(p['lower_bounds']),
            upper_bounds=ren(p['upper_bounds']),
            slopes={to_name[k]: ren(v) for k, v in p['slopes'].items()},
            betas=[None if b is None else ren(b) for b in p['betas']],
            histories=[ren(h) for h in p['histories']], cs=p['cs'], rhs=p['rhs'], sat_solvers=None)
        decisions = [[to_name[d[0]], d[1], d[2]] for d in rec['decisions']]
        res = ab.forward(decisions, params)
        B2 = 2 * len(decisions)
        ref = out['output_lbs']
        assert res.output_lbs.shape == ref.shape and res.output_lbs.device.type == 'cpu'
        assert torch.allclose(res.output_lbs, ref, rtol=1e-5, atol=1e-5 * _scale(ref)), (res.output_lbs - ref).abs().max()
        assert torch.equal(res.output_lbs > out['rhs'], ref > out['rhs'])                 # verdicts
        assert torch.equal(res.rhs.cpu(), out['rhs']) and torch.equal(res.cs.cpu(), out['cs'])
        assert torch.equal(res.objective_ids, out['objective_ids'])
        for k, v in out['lower_bounds'].items():
            if k == 'final':
                continue
            assert torch.equal(res.lower_bounds[to_name[k]], v)
            assert torch.equal(res.upper_bounds[to_name[k]], out['upper_bounds'][k])
        assert torch.isinf(res.upper_bounds[net.final_name]).all()
        for k, v in out['lAs'].items():
            got = res.lAs[to_name[k]]
            assert got.shape == v.shape and got.shape[0] == B2
            assert torch.allclose(got, v, rtol=1e-3, atol=2e-3 * _scale(v))
        for k, v in out['slopes'].items():
            got = res.slopes[to_name[k]][net.final_name]
            assert got.dtype == torch.float16 and got.shape == v['final'].shape
            bad = (got.float() - v['final'].float()).abs() > 4e-3
            assert bad.float().mean() <= 0.01
        assert len(res.betas) == len(res.histories) == B2
        for b, b_ref, h, h_ref in zip(res.betas, out['betas'], res.histories, out['histories']):
            for k in b_ref:
                assert b[to_name[k]].shape == b_ref[k].shape
                assert torch.allclose(b[to_name[k]], b_ref[k], rtol=1e-2, atol=1e-2)
                for a, c in zip(h[to_name[k]], h_ref[k]):
                    assert torch.equal(torch.as_tensor(a).float(), torch.as_tensor(c).float())
