"""Pins the CPU oracle (oracle/crown_oracle.py) to the reference.

* known-answer vectors of the reference's fixed-weight toy net (NS/example/test_model.py:80-108,
  SURVEY.md section 8c),
* golden fixtures recorded from the unmodified reference's BoundedModule.compute_bounds
  (oracle/gen_golden.py): F1 = CROWN pass with reused alpha, F2 = 20-iteration alpha/beta-CROWN.
Tolerance: the north star's 1e-5 relative (fp32), absolute floor 1e-5 on the scale of the terms.
"""
import os

import pytest
import torch

from fixtures import GOLDEN, keyed_inputs, load_fixture
from models import build_model
from neuralsat_b200.graph import activation_indices, preact_indices, trace_module
from oracle import crown_oracle as orc

RTOL = 1e-5
FIXTURES = ['fc_small', 'mnist_fc', 'conv_small', 'resnet_bn_small', 'fc_const', 'oval21_base']
SSHAPE_FIXTURES = ['fc_sigmoid', 'fc_tanh']      # the reference's BaB issues no F1 look-ahead for S-shapes


def close(a, b, rtol=RTOL, atol=1e-5):
    return torch.allclose(a, b, rtol=rtol, atol=atol)


def test_toy_known_answers():
    """l1=[-4,-9,-14] u1=[6,13,20]; l2=[2,-206] u2=[94,148]; lb=[-61,-404,-575.00006];
    lb(C=[1,-1,0]) = -249."""
    model, in_shape = build_model('toy_fixed')
    nodes = trace_module(model, (1, *in_shape))
    pres = preact_indices(nodes)
    x_L = torch.tensor([[-1., -2.]])
    x_U = torch.tensor([[1., 2.]])
    lower = {pres[0]: torch.tensor([[-4., -9., -14.]]), pres[1]: torch.tensor([[2., -206.]])}
    upper = {pres[0]: torch.tensor([[6., 13., 20.]]), pres[1]: torch.tensor([[94., 148.]])}
    C = torch.eye(3).unsqueeze(0)
    lb, _ = orc.crown_pass(nodes, C, x_L, x_U, lower, upper)
    assert close(lb, torch.tensor([[-61., -404., -575.00006]]))
    ub, _ = orc.crown_pass(nodes, -C, x_L, x_U, lower, upper)
    assert close(-ub, torch.tensor([[235., 188., 313.]]))
    lbc, _ = orc.crown_pass(nodes, torch.tensor([[[1., -1., 0.]]]), x_L, x_U, lower, upper)
    assert close(lbc, torch.tensor([[-249.]]))
    # intermediate layer 2 by a pass over the truncated graph (the reference additionally
    # intersects with IBP, which lifts l2[0] from -62 to 2; root bounds are a "next" row)
    sub = nodes[:pres[1] + 1]
    l2, _ = orc.crown_pass(sub, torch.eye(2).unsqueeze(0), x_L, x_U, lower, upper)
    u2, _ = orc.crown_pass(sub, -torch.eye(2).unsqueeze(0), x_L, x_U, lower, upper)
    assert close(l2, torch.tensor([[-62., -206.]])) and close(-u2, torch.tensor([[94., 148.]]))
    # and the generated fixture agrees with the survey's numbers
    fx = torch.load(os.path.join(GOLDEN, 'toy_fixed.pt'), weights_only=False)
    assert close(fx['lb'], torch.tensor([[-61., -404., -575.00006]]))
    assert close(fx['lb_C'], torch.tensor([[-249.]]))


@pytest.mark.parametrize('name', FIXTURES)
def test_f1_pass_matches_reference(name):
    fx, model, nodes = load_fixture(name)
    acts = activation_indices(nodes)
    assert len(fx['f1']) > 0
    for ent in fx['f1']:
        k = keyed_inputs(nodes, ent)
        lb, lA = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'],
                                {r: a[0] for r, a in k['alpha'].items()}, k['alpha_index'], None)
        scale = max(1.0, float(ent['out_lb'].abs().max()))
        assert close(lb, ent['out_lb'], atol=1e-5 * scale), (lb - ent['out_lb']).abs().max()
        for j, r in enumerate(acts):
            assert close(lA[r], ent['out_lA'][j], atol=1e-5 * max(1.0, float(ent['out_lA'][j].abs().max())))


@pytest.mark.parametrize('name', FIXTURES + SSHAPE_FIXTURES)
def test_f2_optimize_matches_reference(name):
    fx, model, nodes = load_fixture(name)
    acts = activation_indices(nodes)
    pres = preact_indices(nodes)
    assert len(fx['f2']) > 0
    for ent in fx['f2']:
        k = keyed_inputs(nodes, ent)
        res = orc.optimize(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], k['alpha'],
                           k['alpha_index'], k['beta'], k['rhs'], iteration=ent['iteration'],
                           lr_alpha=ent['lr_alpha'], lr_beta=ent['lr_beta'], lr_decay=ent['lr_decay'],
                           enable_beta=ent['enable_beta'])
        scale = max(1.0, float(ent['out_lb'].abs().max()))
        # measured: bit-identical on four fixtures, <= 8e-8 * scale on the others
        assert close(res['lb'], ent['out_lb'], rtol=1e-5, atol=1e-5 * scale), \
            (res['lb'] - ent['out_lb']).abs().max()
        # verdict per domain must be identical
        assert torch.equal(res['lb'] > k['rhs'], ent['out_lb'] > k['rhs'])
        for j, r in enumerate(acts):
            # slopes after 20 Adam steps: a step is lr * m / sqrt(v), so a last-bit difference in a tiny gradient moves
            # a slope by up to lr-sized amounts; measured worst case 4.6e-6 (resnet_bn_small), elsewhere 0 - 5e-7
            assert close(res['alpha'][r], ent['out_alpha'][j], rtol=1e-5, atol=2e-5)
            assert close(res['lA'][r], ent['out_lA'][j], rtol=1e-5,
                         atol=1e-5 * max(1.0, float(ent['out_lA'][j].abs().max())))
        for j, p in enumerate(pres):
            assert close(res['beta_val'][p], ent['out_beta_val'][j], rtol=1e-5, atol=1e-5)
