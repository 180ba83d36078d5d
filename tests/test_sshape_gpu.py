"""GPU parity of the sigmoid / tanh relaxation kernels (csrc/crown_sshape.cu)
against reference-recorded fixtures (tests/golden/fc_sigmoid.pt, fc_tanh.pt) and the CPU oracle.
Tolerance: bounds 1e-5 relative (north star), verdicts identical."""
import pytest
import torch

from fixtures import keyed_inputs, load_fixture
from neuralsat_b200.graph import activation_indices, nodes_to, preact_indices
from oracle import crown_oracle as orc
from oracle import sshape_oracle as sso

pytestmark = pytest.mark.gpu
DEV = 'cuda'
FIXTURES = ['fc_sigmoid', 'fc_tanh']


def _scale(t):
    return max(1.0, float(t.abs().max()))


def _lists(nodes, k):
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    lower = [k['lower'][p].to(DEV) for p in pres]
    upper = [k['upper'][p].to(DEV) for p in pres]
    alpha = None if k['alpha'] is None else [k['alpha'][a].to(DEV).contiguous() for a in acts]
    beta = None
    if k['beta'] is not None:
        beta = [{kk: (None if v is None else v.to(DEV).contiguous()) for kk, v in k['beta'][p].items()} for p in pres]
    return lower, upper, alpha, beta


def _plan(nodes):
    from neuralsat_b200 import capi
    return capi.Plan(nodes_to(nodes, DEV))


@pytest.mark.parametrize('name', FIXTURES)
@pytest.mark.parametrize('early_stop', [True, False])
def test_f2_vs_reference(name, early_stop):
    fx, model, nodes = load_fixture(name)
    plan = _plan(nodes)
    for ent in fx['f2']:
        k = keyed_inputs(nodes, ent)
        lower, upper, alpha, beta = _lists(nodes, k)
        lb, lA, n_iter = plan.optimize(k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV), lower, upper, alpha, None,
                                       beta, k['rhs'].to(DEV), iteration=ent['iteration'], lr_alpha=ent['lr_alpha'],
                                       lr_beta=ent['lr_beta'], lr_decay=ent['lr_decay'],
                                       enable_beta=ent['enable_beta'], early_stop=early_stop)
        ref = ent['out_lb']
        assert torch.allclose(lb.cpu(), ref, rtol=1e-5, atol=1e-5 * _scale(ref)), (lb.cpu() - ref).abs().max()
        assert torch.equal(lb.cpu() > k['rhs'], ref > k['rhs'])
        n_bad = n_all = 0
        for j in range(len(lA)):
            r = ent['out_lA'][j]
            assert torch.allclose(lA[j].cpu(), r, rtol=1e-3, atol=2e-3 * _scale(r))
            bad = (alpha[j].cpu() - ent['out_alpha'][j]).abs() > 2e-3 + 1e-3 * ent['out_alpha'][j].abs()
            n_bad += int(bad.sum())
            n_all += bad.numel()
        assert n_bad <= 0.01 * n_all, (n_bad, n_all)
        if beta is not None:
            for j, bt in enumerate(beta):
                if bt['val'].numel():
                    assert torch.allclose(bt['val'].cpu(), ent['out_beta_val'][j], rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize('name', FIXTURES)
def test_pass_and_grad_vs_oracle(name):
    fx, model, nodes = load_fixture(name)
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    plan = _plan(nodes)
    ent = fx['f2'][-1]
    k = keyed_inputs(nodes, ent)
    a_par = {r: a.clone().requires_grad_() for r, a in k['alpha'].items()}
    b_par = {p: b['val'].clone().requires_grad_() for p, b in k['beta'].items()}
    beta_o = {p: dict(b, val=b_par[p]) for p, b in k['beta'].items()}
    lb_o, lA_o = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], a_par, k['alpha_index'], beta_o)
    lb_o.sum().backward()
    lower, upper, alpha, beta = _lists(nodes, k)
    lb, lA, ga, gb = plan.crown_grad(k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV), lower, upper, alpha, None, beta)
    assert torch.allclose(lb.cpu(), lb_o.detach(), rtol=1e-5, atol=1e-5 * _scale(lb_o)), (lb.cpu() - lb_o).abs().max()
    for j, a in enumerate(acts):
        # the pass clips the tangent points in place, like the reference (operators/tanh.py:191-198)
        assert torch.allclose(alpha[j].cpu(), a_par[a].detach(), rtol=0, atol=1e-6)
        assert torch.allclose(lA[j].cpu(), lA_o[a].detach(), rtol=1e-5, atol=1e-5 * _scale(lA_o[a]))
        ref = a_par[a].grad
        assert torch.allclose(ga[j].cpu(), ref, rtol=1e-4, atol=1e-5 * _scale(ref)), (j, (ga[j].cpu() - ref).abs().max())
        assert ref[0::2].abs().sum() > 0 and float(ga[j][1::2].abs().sum()) == 0.0
    for j, p in enumerate(pres):
        if gb[j] is None:
            continue
        ref = b_par[p].grad
        assert torch.allclose(gb[j].cpu(), ref, rtol=1e-4, atol=1e-5 * _scale(ref))
    # plain CROWN lines (no tangent-point parameters): middle-point / table tangents
    lb2, _ = plan.crown_pass(k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV), lower, upper, None, None, None)
    lb2_o, _ = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'])
    assert torch.allclose(lb2.cpu(), lb2_o, rtol=1e-5, atol=1e-5 * _scale(lb2_o)), (lb2.cpu() - lb2_o).abs().max()


@pytest.mark.parametrize('name,Bd,S', [('fc_sigmoid', 130, 1), ('fc_tanh', 67, 3)])
def test_synthetic_vs_oracle(name, Bd, S):
    """Seeded synthetic batch: wide interval bounds so that all three cases (l>=0, u<=0, crossing) and
    the direct-line branches occur; S > 1 with per-spec tangent points (S1 = S); 4 optimiser iterations."""
    fx, model, nodes = load_fixture(name)
    op = 'sigmoid' if name == 'fc_sigmoid' else 'tanh'
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    g = torch.Generator().manual_seed(3)
    x0 = torch.rand(Bd, *nodes[0]['shape'], generator=g)
    eps = 0.01 + 0.3 * torch.rand(Bd, 1, generator=g)
    x_L, x_U = x0 - eps, x0 + eps
    pre = orc.interval_bounds(nodes, x_L, x_U)
    lower = {p: pre[p][0].clone() for p in pres}
    upper = {p: pre[p][1].clone() for p in pres}
    n_out = nodes[-1]['shape'][0]
    C = torch.randn(Bd, S, n_out, generator=g)
    alpha = {a: sso.init_alpha(op, lower[p], upper[p], S1=S) + 0.05 * torch.randn(8, S, Bd, *lower[p].shape[1:], generator=g)
             for a, p in zip(acts, pres)}
    beta = {}
    for p in pres:
        n = lower[p][0].numel()
        J = 3
        loc = torch.randint(0, n, (Bd, J), generator=g)
        sign = (torch.randint(0, 2, (Bd, J), generator=g) * 2 - 1).float()
        point = torch.gather((lower[p] + upper[p]) / 2, 1, loc)
        beta[p] = {'val': torch.rand(Bd, J, generator=g) * 0.05, 'loc': loc, 'sign': sign, 'bias': point}
    k = dict(C=C, x_L=x_L, x_U=x_U, lower=lower, upper=upper, alpha=alpha, alpha_index={a: None for a in acts},
             beta=beta, rhs=torch.zeros(Bd, S))
    plan = _plan(nodes)
    res = orc.optimize(nodes, C, x_L, x_U, lower, upper, alpha, k['alpha_index'], beta, k['rhs'], iteration=4)
    lo, up, al, bt = _lists(nodes, k)
    lb, lA, _ = plan.optimize(C.to(DEV), x_L.to(DEV), x_U.to(DEV), lo, up, al, None, bt, k['rhs'].to(DEV), iteration=4)
    print(f'[{name}] max |lb - oracle| = {float((lb.cpu() - res["lb"]).abs().max()):.3e} (scale {_scale(res["lb"]):.2f})')
    assert torch.allclose(lb.cpu(), res['lb'], rtol=1e-5, atol=2e-5 * _scale(res['lb'])), (lb.cpu() - res['lb']).abs().max()
    # The optimised tangent points are compared element-wise, but a handful may legitimately differ: Adam divides the
    # gradient by its own magnitude, so an element whose gradient is an exact 0 in one summation order and round-off in
    # another (k_tc_linear adds the bias terms of different column tiles with float atomics: the order is run-dependent,
    # the test fails about one run in three without this) moves by +-lr.  Such elements do not move lb (checked above
    # to 1e-5), so: at most 2 % outliers, everything else to 1e-3.  Measured over 30 runs on a B200: 26 runs with 0
    # elements off (max 3e-6), 4 runs with 151 - 168 of 38592 elements (0.4 %) off by up to 0.18 in one of two
    # repeating patterns (= the few possible orders of the atomic adds), lb identical to 2e-6 in all of them.
    for j, a in enumerate(acts):
        got, ref = al[j].cpu(), res['alpha'][a]
        bad = ~torch.isclose(got, ref, rtol=1e-3, atol=2e-3)
        print(f'[{name}] alpha {j}: {int(bad.sum())} of {bad.numel()} elements off, max {float((got - ref).abs().max()):.3e}')
        assert bad.float().mean().item() <= 2e-2, (int(bad.sum()), bad.numel(), (got - ref).abs().max())
