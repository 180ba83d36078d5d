"""GPU parity at the REAL architectures of BASELINE.json configs 3, 4, 5: the CUDA path against the
CPU oracle on seeded synthetic sub-domain batches, at batch sizes the oracle finishes in seconds.

Covers what the toy fixtures cannot: convolutions with 16-128 channels on 32x32 ... 2x2 maps (the tensor-core
implicit-GEMM kernels, and with CROWN_B200_DISABLE_CONV_TC=1 the register-tiled / direct SIMT kernels with their
shared-memory size gates), kernel 4 stride 2, kernel 3 stride 2 without padding, 1x1 stride-2 shortcuts, residual Add
fan-in with and without a ReLU behind it, explicit BatchNorm, and beta record lists longer than the chain kernels'
shared-memory table (CHAIN_JMAX = 32)."""
import pytest
import torch

from neuralsat_b200.graph import activation_indices, nodes_to, preact_indices, trace_module
from oracle import crown_oracle as orc

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _scale(t):
    return max(1.0, float(t.abs().max()))


def _problem(workload, Bd, seed, max_splits=8):
    from neuralsat_b200 import synth
    fold = workload.endswith(':folded')              # BatchNorm merged into the convolutions (what the ONNX loader does)
    workload = workload.split(':')[0]
    wl = synth.WORKLOADS[workload]
    nodes = synth.build_nodes(workload, seed=0, fold_bn=fold)
    b = synth.make_batch(nodes, Bd, wl['eps'], seed=seed, device='cpu', max_splits=max_splits,
                         bounds=wl.get('bounds', 'ibp'))
    g = torch.Generator().manual_seed(seed + 7)
    for bt in b['beta']:                                  # non-zero multipliers so that beta shows in the bounds
        bt['val'] = torch.rand(bt['val'].shape, generator=g) * 0.05 * (bt['sign'] != 0)
    # per-domain boxes and margin rows, so that rows differ in more than their splits
    n_out = b['C'].shape[-1]
    b['C'] = torch.randn(Bd, 1, n_out, generator=g)
    shrink = 0.5 + 0.5 * torch.rand(Bd, *[1] * (b['x_L'].dim() - 1), generator=g)
    c, r = (b['x_L'] + b['x_U']) / 2, (b['x_U'] - b['x_L']) / 2
    b['x_L'], b['x_U'] = (c - r * shrink).contiguous(), (c + r * shrink).contiguous()
    return nodes, b


def _keyed(nodes, b):
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    return dict(C=b['C'], x_L=b['x_L'], x_U=b['x_U'],
                lower={p: b['lower'][k] for k, p in enumerate(pres)},
                upper={p: b['upper'][k] for k, p in enumerate(pres)},
                alpha={a: b['alpha'][k] for k, a in enumerate(acts)},
                alpha_index={a: None for a in acts},
                beta={p: b['beta'][k] for k, p in enumerate(pres)})


def _dev(b):
    return dict(C=b['C'].to(DEV), x_L=b['x_L'].to(DEV), x_U=b['x_U'].to(DEV),
                lower=[t.to(DEV) for t in b['lower']], upper=[t.to(DEV) for t in b['upper']],
                alpha=[t.to(DEV).contiguous() for t in b['alpha']],
                beta=[{k: (None if v is None else v.to(DEV).contiguous()) for k, v in bt.items()} for bt in b['beta']])


def _grad_close(got, ref, what, max_bad=5e-3):
    """Gradients at these sizes: 1e-4 relative to the largest entry of the tensor for all but a few entries in a thousand.
    The alpha / beta gradient is a SUB-gradient: every neuron whose coefficient A is within round-off of zero may take
    either line of the sign-split multiply (the reference's A >= 0 rule decides on ITS rounding of A), and one such
    neuron changes the gradient of everything downstream of it.  Deep networks and long beta lists (dozens of +-beta
    terms added to A) produce a handful of such ties per batch; the bounds themselves are compared to 1e-5."""
    tol = 1e-4 * max(float(ref.abs().max()), 1e-6) + 1e-4 * ref.abs()
    bad = (got - ref).abs() > tol
    assert bad.float().mean() <= max_bad, (what, int(bad.sum()), bad.numel(), float((got - ref).abs().max()), float(ref.abs().max()))
    assert float((got - ref).abs().max()) <= 0.05 * float(ref.abs().max())


CASES = [('oval21_base', 64), ('sri_resnet_a', 48), ('cifar10_2_255', 24), ('cifar100_resnet_medium', 12),
         ('cifar100_resnet_medium:folded', 12)]


@pytest.fixture(params=['conv_tc', 'conv_simt'])
def conv_path(request, monkeypatch):
    """'conv_tc': convolutions on the tcgen05 implicit-GEMM kernels (default); 'conv_simt': the fp32 SIMT kernels."""
    monkeypatch.setenv('CROWN_B200_DISABLE_CONV_TC', '1' if request.param == 'conv_simt' else '0')
    monkeypatch.setenv('CROWN_B200_CONV_AUTOTUNE', '0')        # parity of the tensor-core kernel on EVERY layer
    return request.param


@pytest.mark.parametrize('workload,Bd', CASES)
def test_pass_and_gradient_vs_oracle(workload, Bd, conv_path):
    from neuralsat_b200 import capi
    nodes, b = _problem(workload, Bd, seed=3)
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    plan = capi.Plan(nodes_to(nodes, DEV))
    if conv_path == 'conv_tc':
        assert plan.conv_tc > 0, 'the tensor-core convolution path must be the one that runs'
    k = _keyed(nodes, b)
    a_par = {r: a.clone().requires_grad_() for r, a in k['alpha'].items()}
    b_par = {p: bt['val'].clone().requires_grad_() for p, bt in k['beta'].items()}
    beta_o = {p: dict(bt, val=b_par[p]) for p, bt in k['beta'].items()}
    lb_o, lA_o = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'],
                                {r: a[0] for r, a in a_par.items()}, k['alpha_index'], beta_o)
    lb_o.sum().backward()
    d = _dev(b)
    lb, lA, ga, gb = plan.crown_grad(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None, d['beta'])
    assert torch.allclose(lb.cpu(), lb_o.detach(), rtol=1e-5, atol=1e-5 * _scale(lb_o)), (lb.cpu() - lb_o).abs().max()
    for j, a in enumerate(acts):
        ref = lA_o[a].detach()
        assert torch.allclose(lA[j].cpu(), ref, rtol=1e-5, atol=1e-5 * _scale(ref)), (j, (lA[j].cpu() - ref).abs().max())
        ref = a_par[a].grad[0]
        _grad_close(ga[j].cpu(), ref, f'grad_alpha[{j}]')
    for j, p in enumerate(pres):
        if gb[j] is not None:
            _grad_close(gb[j].cpu(), b_par[p].grad, f'grad_beta[{j}]')
    # F1 form: adaptive slopes, no beta
    lb2, _ = plan.crown_pass(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], None, None, None, want_lA=False)
    lb2_o, _ = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'])
    assert torch.allclose(lb2.cpu(), lb2_o, rtol=1e-5, atol=1e-5 * _scale(lb2_o)), (lb2.cpu() - lb2_o).abs().max()


@pytest.mark.parametrize('workload,Bd', CASES)
def test_short_optimisation_vs_oracle(workload, Bd, conv_path):
    from neuralsat_b200 import capi
    nodes, b = _problem(workload, Bd, seed=5)
    plan = capi.Plan(nodes_to(nodes, DEV))
    k = _keyed(nodes, b)
    rhs = torch.zeros(Bd, 1)
    res = orc.optimize(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], k['alpha'], k['alpha_index'],
                       k['beta'], rhs, iteration=5)
    d = _dev(b)
    lb, lA, _ = plan.optimize(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None, d['beta'],
                              rhs.to(DEV), iteration=5)
    assert torch.allclose(lb.cpu(), res['lb'], rtol=1e-4, atol=1e-4 * _scale(res['lb'])), (lb.cpu() - res['lb']).abs().max()
    assert torch.equal(lb.cpu() > rhs, res['lb'] > rhs) or ((lb.cpu() - res['lb']).abs() < 1e-5 * _scale(res['lb'])).all()


def test_tinyimagenet_resnet_pass_vs_oracle():
    """BASELINE.json configs[4], the 56x56 variant (27x27 / 14x14 / 7x7 maps): one pass, 6 sub-domains."""
    from neuralsat_b200 import capi
    nodes, b = _problem('tinyimagenet_resnet_medium', 6, seed=2)
    plan = capi.Plan(nodes_to(nodes, DEV))
    k = _keyed(nodes, b)
    lb_o, _ = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'],
                             {r: a[0] for r, a in k['alpha'].items()}, k['alpha_index'], k['beta'])
    d = _dev(b)
    lb, _ = plan.crown_pass(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None, d['beta'], want_lA=False)
    assert torch.allclose(lb.cpu(), lb_o, rtol=1e-5, atol=1e-5 * _scale(lb_o)), (lb.cpu() - lb_o).abs().max()


@pytest.mark.parametrize('J', [33, 48, 56, 64, 80, 128])
def test_beta_lists_longer_than_the_chain_table(J):
    """More than CHAIN_JMAX = 32 records per row and layer: the whole-network kernels hand over to the per-layer
    tensor-core kernels (crown_api.cu:chain_applies); pass, gradient and a short optimisation against the oracle."""
    from neuralsat_b200 import capi
    nodes, b = _problem('mnistfc_256x4', 96, seed=11, max_splits=4 * J)
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    g = torch.Generator().manual_seed(J)
    for kk, bt in enumerate(b['beta']):                  # exactly J records on every layer, all live, on unstable neurons
        n = b['lower'][kk][0].numel()
        bt['loc'] = torch.randint(0, n, (96, J), generator=g)
        bt['sign'] = (torch.randint(0, 2, (96, J), generator=g) * 2 - 1).float()
        bt['val'] = torch.rand(96, J, generator=g) * 0.05
        bt['bias'] = None
    plan = capi.Plan(nodes_to(nodes, DEV))
    assert plan.chain
    k = _keyed(nodes, b)
    a_par = {r: a.clone().requires_grad_() for r, a in k['alpha'].items()}
    b_par = {p: bt['val'].clone().requires_grad_() for p, bt in k['beta'].items()}
    beta_o = {p: dict(bt, val=b_par[p]) for p, bt in k['beta'].items()}
    lb_o, _ = orc.crown_pass(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'],
                             {r: a[0] for r, a in a_par.items()}, k['alpha_index'], beta_o)
    lb_o.sum().backward()
    d = _dev(b)
    lb, lA, ga, gb = plan.crown_grad(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None, d['beta'])
    assert torch.allclose(lb.cpu(), lb_o.detach(), rtol=1e-5, atol=1e-5 * _scale(lb_o)), (lb.cpu() - lb_o).abs().max()
    for j, a in enumerate(acts):
        _grad_close(ga[j].cpu(), a_par[a].grad[0], f'grad_alpha[{j}]', max_bad=2.5e-2)
    for j, p in enumerate(pres):
        # one tie (see _grad_close) moves every beta gradient of its row: allow two of the 96 rows
        _grad_close(gb[j].cpu(), b_par[p].grad, f'grad_beta[{j}]', max_bad=2.5e-2)
    rhs = torch.zeros(96, 1)
    res = orc.optimize(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], k['alpha'], k['alpha_index'],
                       k['beta'], rhs, iteration=4)
    lb3, _, _ = plan.optimize(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None, d['beta'],
                              rhs.to(DEV), iteration=4)
    assert torch.allclose(lb3.cpu(), res['lb'], rtol=1e-4, atol=1e-4 * _scale(res['lb']))


def test_conv_choices_are_reported_and_replayed(monkeypatch):
    """cb_plan_conv_choices / CROWN_B200_CONV_CHOICES: a plan reports which kernel each (convolution, direction) runs,
    and a second plan replays that string instead of timing the kernels (profiling runs depend on it); a string of the
    wrong length is ignored."""
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to
    nodes = nodes_to(synth.build_nodes('oval21_base', seed=0), 'cuda')
    n_conv = sum(1 for nd in nodes if nd['op'] == 'conv2d')
    monkeypatch.setenv('CROWN_B200_CONV_AUTOTUNE', '0')
    monkeypatch.delenv('CROWN_B200_CONV_CHOICES', raising=False)
    all_tc = capi.Plan(nodes)
    assert all_tc.conv_choices == 'T' * (2 * n_conv) and all_tc.conv_tc == 2 * n_conv
    want = 'ST' * n_conv
    monkeypatch.setenv('CROWN_B200_CONV_CHOICES', want)
    replay = capi.Plan(nodes)
    assert replay.conv_choices == want and replay.conv_tc == n_conv
    monkeypatch.setenv('CROWN_B200_CONV_CHOICES', 'S')             # wrong length: not applied
    assert capi.Plan(nodes).conv_choices == 'T' * (2 * n_conv)
    # the replayed plan computes the same bounds
    wl = synth.WORKLOADS['oval21_base']
    b = synth.make_batch(nodes, 16, wl['eps'], seed=3, device='cuda', bounds=wl.get('bounds', 'ibp'))
    args = (b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'])
    lb1, _ = all_tc.crown_pass(*args, want_lA=False)
    lb2, _ = replay.crown_pass(*args, want_lA=False)
    assert torch.allclose(lb1, lb2, rtol=1e-5, atol=1e-5 * max(1.0, float(lb1.abs().max())))


@pytest.mark.parametrize('switch', ['CROWN_B200_DISABLE_BETA_IN_RELU', 'CROWN_B200_DISABLE_ADAM_IN_GRAD',
                                    'CROWN_B200_DISABLE_SEED_IN_CONCRETIZE'])
def test_folded_kernels_match_their_stand_alone_form(switch, monkeypatch):
    """The split constraints ride in relu_bwd / relu_grad, the Adam step of the slopes in relu_grad and the seed of the
    gradient sweep in the concretize launch; with the switch set the stand-alone beta_scatter / beta_grad / k_adam /
    grad_init launches run instead.  Same arithmetic per element: the 5-step
    trajectories must agree to rounding (the beta bias joins a different partial sum)."""
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to
    nodes = nodes_to(synth.build_nodes('sri_resnet_a', seed=0), 'cuda')
    wl = synth.WORKLOADS['sri_resnet_a']
    plan = capi.Plan(nodes)

    def run():
        b = synth.make_batch(nodes, 24, wl['eps'], seed=5, device='cuda', bounds=wl.get('bounds', 'ibp'))
        n0 = capi.launch_count()
        lb, lA, _ = plan.optimize(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'], None,
                                  iteration=5, early_stop=False)
        return lb, [a.clone() for a in b['alpha']], [bt['val'].clone() for bt in b['beta']], capi.launch_count() - n0

    monkeypatch.delenv(switch, raising=False)
    lb1, al1, be1, n1 = run()
    monkeypatch.setenv(switch, '1')
    lb2, al2, be2, n2 = run()
    # the stand-alone beta launches are really there (k_adam is launched either way: it skips the folded tensors inside)
    assert n2 == n1 if 'ADAM' in switch else n2 > n1
    assert torch.allclose(lb1, lb2, rtol=1e-5, atol=1e-5 * max(1.0, float(lb1.abs().max())))
    for x, y in zip(al1 + be1, al2 + be2):
        assert torch.allclose(x, y, rtol=1e-4, atol=1e-4)
