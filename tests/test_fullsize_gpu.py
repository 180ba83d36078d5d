"""GPU, BASELINE.json sizes: size-independent properties of the CUDA path at the batch sizes the bench uses (the CPU
oracle would take minutes there).

* batch independence: every sub-domain is an independent unit, so bounding a batch must give, row for
  row, the bits obtained by bounding its slices separately (also the invariant the multi-GPU sharding relies on);
* keep-best monotonicity: the optimised bound is never below the first pass' bound (AL/optimized_bounds.py:180-204);
* soundness: with valid intermediate bounds and no splits, lb is a lower bound of C.f(x) at sampled points of the box;
* determinism: two runs give identical bits on the whole-network kernels (fixed summation order); the per-layer
  tensor-core kernels accumulate the bias terms of different column tiles with float atomics, so the conv config is
  compared to fp32 round-off (1e-5 relative after 20 Adam iterations) instead.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _setup(workload, Bd, seed=0, max_splits=16, bound_scale=0.25):
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to, trace_module
    wl = synth.WORKLOADS[workload]
    net = synth.build_network(workload, seed=0)
    nodes = trace_module(net, (1, *wl['in_shape']))
    plan = capi.Plan(nodes_to(nodes, DEV))
    b = synth.make_batch(nodes, Bd, wl['eps'], seed=seed, device=DEV, max_splits=max_splits, bound_scale=bound_scale)
    return net.to(DEV), nodes, plan, b


def _slice(b, lo, hi):
    return {'C': b['C'][lo:hi].contiguous(), 'x_L': b['x_L'][lo:hi].contiguous(), 'x_U': b['x_U'][lo:hi].contiguous(),
            'lower': [t[lo:hi].contiguous() for t in b['lower']], 'upper': [t[lo:hi].contiguous() for t in b['upper']],
            'alpha': [t[:, :, lo:hi].contiguous() for t in b['alpha']],
            'beta': [{k: (None if v is None else v[lo:hi].contiguous()) for k, v in bt.items()} for bt in b['beta']]}


def _f2(plan, b, iteration=20):
    alpha = [a.clone() for a in b['alpha']]
    beta = [dict(bt, val=bt['val'].clone()) for bt in b['beta']]
    lb, lA, _ = plan.optimize(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], alpha, None, beta, None,
                              iteration=iteration, early_stop=False, early_stop_patience=10 ** 6, want_lA=True)
    return lb, lA, alpha, beta


@pytest.mark.parametrize('workload,Bd,cut', [('mnistfc_256x4', 8192, 3072), ('oval21_base', 4096, 1000)])
def test_batch_independence_and_determinism(workload, Bd, cut):
    _, nodes, plan, b = _setup(workload, Bd)
    exact = plan.chain

    def same(x, y):
        if exact:
            return torch.equal(x, y)
        return torch.allclose(x, y, rtol=1e-5, atol=1e-5 * max(1.0, float(y.abs().max())))

    lb, lA, alpha, beta = _f2(plan, b)
    lb2, lA2, alpha2, _ = _f2(plan, b)
    assert same(lb, lb2)                                                                          # determinism
    if exact:
        assert all(torch.equal(x, y) for x, y in zip(alpha, alpha2))
    assert torch.isfinite(lb).all()
    parts = [_f2(plan, _slice(b, 0, cut)), _f2(plan, _slice(b, cut, Bd))]
    lb_cat = torch.cat([p[0] for p in parts])
    assert same(lb, lb_cat), (lb - lb_cat).abs().max()
    if exact:
        for k in range(len(alpha)):
            assert torch.equal(alpha[k], torch.cat([p[2][k] for p in parts], dim=2))
            assert torch.equal(lA[k], torch.cat([p[1][k] for p in parts], dim=1))
        for k in range(len(beta)):
            assert torch.equal(beta[k]['val'], torch.cat([p[3][k]['val'] for p in parts]))


@pytest.mark.parametrize('workload,Bd', [('mnistfc_256x4', 8192), ('oval21_base', 4096)])
def test_keep_best_is_monotone(workload, Bd):
    _, nodes, plan, b = _setup(workload, Bd, seed=1)
    lb1, _ = plan.crown_pass(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'], want_lA=False)
    lb20, _, _, _ = _f2(plan, b)
    assert (lb20 >= lb1).all()
    assert (lb20 > lb1 + 1e-6).float().mean() > 0.5          # and the optimiser does improve most sub-domains


@pytest.mark.parametrize('workload,Bd', [('mnistfc_256x4', 8192), ('oval21_base', 4096)])
def test_lower_bound_is_sound_without_splits(workload, Bd):
    net, nodes, plan, b = _setup(workload, Bd, seed=2, max_splits=0, bound_scale=1.0)      # plain interval bounds: sound
    beta = None
    alpha = [a.clone() for a in b['alpha']]
    lb, _, _ = plan.optimize(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], alpha, None, beta, None,
                             iteration=10, early_stop=False, early_stop_patience=10 ** 6, want_lA=False)
    g = torch.Generator(device=DEV).manual_seed(0)
    worst = torch.full_like(lb, float('inf'))
    with torch.no_grad():
        for _ in range(8):
            t = torch.rand(b['x_L'].shape, device=DEV, generator=g)
            x = b['x_L'] + t * (b['x_U'] - b['x_L'])
            y = net(x.view(Bd, *nodes[0]['shape']))
            worst = torch.minimum(worst, torch.einsum('bsn,bn->bs', b['C'], y))
        for x in (b['x_L'], b['x_U']):
            y = net(x.view(Bd, *nodes[0]['shape']))
            worst = torch.minimum(worst, torch.einsum('bsn,bn->bs', b['C'], y))
    slack = 1e-4 * worst.abs().clamp(min=1.0)
    assert (lb <= worst + slack).all(), (lb - worst).max()
