"""GPU: the host-buffer pipeline (neuralsat_b200.pipeline.HostPipeline) returns exactly what the blocking
call returns, for consecutive batches of different sizes (slot reuse, reallocation, fp16 slopes)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _host(b, half_alpha=False):
    pin = lambda t: t.cpu().pin_memory()
    return {'C': pin(b['C']), 'x_L': pin(b['x_L']), 'x_U': pin(b['x_U']),
            'lower': [pin(t) for t in b['lower']], 'upper': [pin(t) for t in b['upper']],
            'alpha': [pin(t.half() if half_alpha else t) for t in b['alpha']],
            'beta': [{k: (None if v is None else pin(v)) for k, v in bt.items()} for bt in b['beta']]}


@pytest.mark.parametrize('half_alpha', [False, True])
def test_pipeline_matches_blocking_call(half_alpha):
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to, trace_module
    from neuralsat_b200.pipeline import HostPipeline
    net = synth.build_network('mnistfc_256x4', seed=0)
    nodes = trace_module(net, (1, 1, 28, 28))
    plan = capi.Plan(nodes_to(nodes, DEV))
    batches = [synth.make_batch(nodes, bd, 0.02, seed=s, device=DEV) for s, bd in enumerate([96, 96, 70, 96, 130])]
    hosts = [_host(b, half_alpha) for b in batches]       # fp16 slopes on the host: widened on the device (exact)
    kw = dict(iteration=6, early_stop=False, early_stop_patience=10 ** 6, want_lA=True)
    ref = []
    for b in batches:
        alpha = [a.clone() for a in b['alpha']]
        beta = [dict(bt, val=bt['val'].clone()) for bt in b['beta']]
        lb, lA, _ = plan.optimize(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], alpha, None, beta, None, **kw)
        ref.append({'lb': lb.cpu(), 'lA': [t.cpu() for t in lA], 'alpha': [a.half().cpu() for a in alpha],
                    'beta': [bt['val'].cpu() for bt in beta]})
    pipe = HostPipeline(plan, depth=2, **kw)
    tickets, got = [], []
    for i, h in enumerate(hosts):
        tickets.append(pipe.submit(h))
        if i >= 1:
            r = pipe.result(tickets[i - 1])
            got.append({k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in r.items()})
    r = pipe.result(tickets[-1])
    got.append({k: ([t.clone() for t in v] if isinstance(v, list) else v.clone()) for k, v in r.items()})
    pipe.drain()
    torch.cuda.synchronize()
    for g, e in zip(got, ref):
        assert torch.equal(g['lb'], e['lb'])
        for a, b in zip(g['lA'], e['lA']):
            assert torch.equal(a, b)
        for a, b in zip(g['alpha'], e['alpha']):
            assert a.dtype == torch.float16 and torch.equal(a, b)
        for a, b in zip(g['beta'], e['beta']):
            assert torch.equal(a, b)
    assert pipe.total_in > 0 and pipe.total_out > 0
