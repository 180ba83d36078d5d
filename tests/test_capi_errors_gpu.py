"""GPU: error behaviour of the C-ABI (include/crown_b200.h) as seen through the binding: status codes become
RuntimeErrors, allocation failures carry the exact message the reference's OOM back-off matches
(NS/util/misc/torch_cuda_memory.py:58-61: a RuntimeError with one argument containing "CUDA out of memory.")."""
import ctypes as C

import pytest
import torch

from fixtures import keyed_inputs, load_fixture
from neuralsat_b200.graph import activation_indices, nodes_to, preact_indices

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _problem(plan, nodes, ent):
    k = keyed_inputs(nodes, ent)
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    lower = [k['lower'][p].to(DEV) for p in pres]
    upper = [k['upper'][p].to(DEV) for p in pres]
    alpha = [k['alpha'][a].to(DEV).contiguous() for a in acts]
    from neuralsat_b200 import capi
    pos = [capi.alpha_pos_from_index(k['alpha_index'][a], int(k['lower'][p][0].numel()), DEV) for a, p in zip(acts, pres)]
    return k, lower, upper, alpha, pos


def is_cuda_oom(e):            # the reference's matcher, restated
    return isinstance(e, RuntimeError) and len(e.args) == 1 and 'CUDA out of memory.' in e.args[0]


def test_status_codes_and_messages():
    from neuralsat_b200 import capi
    fx, model, nodes = load_fixture('fc_small')
    plan = capi.Plan(nodes_to(nodes, DEV))
    L = capi.lib()
    k, lower, upper, alpha, pos = _problem(plan, nodes, fx['f1'][0])
    Cm, xl, xu = k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV)
    lb, lA = plan._alloc_out(int(Cm.shape[0]), int(Cm.shape[1]), True)
    pr, keep = plan._problem(Cm, xl, xu, lower, upper, alpha, pos, None, lb, lA)
    need = L.cb_workspace_bytes(plan.handle, pr.Bd, pr.S, 0, C.byref(pr))
    assert need > 0
    ws = torch.empty(need, dtype=torch.uint8, device=DEV)
    stream = torch.cuda.current_stream().cuda_stream
    assert L.cb_crown_pass(plan.handle, C.byref(pr), ws.data_ptr(), ws.numel(), stream) == 0
    # workspace too small -> CB_ERR_WORKSPACE (4), nothing launched
    n0 = capi.launch_count()
    assert L.cb_crown_pass(plan.handle, C.byref(pr), ws.data_ptr(), 1024, stream) == 4
    assert b'workspace' in L.cb_last_error() and capi.launch_count() == n0
    # null workspace / null problem / bad sizes -> CB_ERR_WORKSPACE / CB_ERR_ARG (1)
    assert L.cb_crown_pass(plan.handle, C.byref(pr), None, 0, stream) == 4
    assert L.cb_crown_pass(plan.handle, None, ws.data_ptr(), ws.numel(), stream) == 1
    pr.Bd = 0
    assert L.cb_crown_pass(plan.handle, C.byref(pr), ws.data_ptr(), ws.numel(), stream) == 1
    with pytest.raises(RuntimeError, match='crown_b200 error 1'):
        capi._check(1)
    # optimiser options are validated
    pr.Bd = int(Cm.shape[0])
    opt = capi.CbOpt()
    opt.iteration = 0
    assert L.cb_optimize(plan.handle, C.byref(pr), C.byref(opt), ws.data_ptr(), ws.numel(), stream, None) == 1
    torch.cuda.synchronize()


def test_unsupported_graph_is_rejected():
    from neuralsat_b200 import capi
    bad = [{'op': 'input', 'in': [], 'shape': (4,)},
           {'op': 'linear', 'in': [0], 'shape': (3,), 'weight': torch.zeros(3, 4, device=DEV), 'bias': None},
           {'op': 'add', 'in': [0, 1], 'shape': (3,)}]                   # broadcasting add: shapes differ
    with pytest.raises(RuntimeError, match='broadcast'):
        capi.Plan(bad)


def test_allocation_failure_matches_the_reference_oom_matcher():
    from neuralsat_b200 import capi
    fx, model, nodes = load_fixture('fc_small')
    plan = capi.Plan(nodes_to(nodes, DEV))
    with pytest.raises(RuntimeError) as ei:
        plan._workspace(1 << 46)                       # 64 TiB
    assert is_cuda_oom(ei.value), ei.value.args
    try:
        capi._check(capi.CB_ERR_OOM)
    except RuntimeError as e:
        assert is_cuda_oom(e)
    # the plan stays usable after the failed allocation
    k, lower, upper, alpha, pos = _problem(plan, nodes, fx['f1'][0])
    lb, _ = plan.crown_pass(k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV), lower, upper, alpha, pos, None)
    ref = fx['f1'][0]['out_lb']
    assert torch.allclose(lb.cpu(), ref, rtol=1e-5, atol=1e-5 * max(1.0, float(ref.abs().max())))


def test_malformed_problems_are_rejected_before_any_launch():
    """Shapes, devices and index ranges of a problem are checked in the binding (a bad beta `loc` would otherwise index
    shared memory inside the chain kernel)."""
    from neuralsat_b200 import capi
    fx, model, nodes = load_fixture('fc_small')
    plan = capi.Plan(nodes_to(nodes, DEV))
    ent = fx['f2'][0] if 'f2' in fx else fx['f1'][0]
    k, lower, upper, alpha, pos = _problem(plan, nodes, ent)
    Cm, xl, xu = k['C'].to(DEV), k['x_L'].to(DEV), k['x_U'].to(DEV)
    Bd, S = int(Cm.shape[0]), int(Cm.shape[1])
    lb, lA = plan._alloc_out(Bd, S, True)
    n0 = capi.launch_count()
    ok = lambda **kw: plan._problem(kw.get('C', Cm), kw.get('x_L', xl), kw.get('x_U', xu), kw.get('lower', lower),
                                    kw.get('upper', upper), kw.get('alpha', alpha), kw.get('pos', pos), kw.get('beta'), lb, lA)
    ok()
    with pytest.raises(ValueError, match='C must be'):
        ok(C=Cm[..., :-1].contiguous())
    with pytest.raises(ValueError, match='x_L / x_U'):
        ok(x_L=xl[:, :-1].contiguous())
    with pytest.raises(ValueError, match='intermediate bound'):
        ok(lower=lower[:-1])
    with pytest.raises(ValueError, match='batch dimension'):
        ok(alpha=[a[:, :, :-1].contiguous() for a in alpha])
    with pytest.raises(ValueError, match='expected'):
        ok(alpha=[a.reshape(-1) for a in alpha])
    J = 3
    n0k = plan.act_numel[0]
    good = {'val': torch.zeros(Bd, J, device=DEV), 'loc': torch.zeros(Bd, J, dtype=torch.int64, device=DEV),
            'sign': torch.ones(Bd, J, device=DEV)}
    beta = [good] + [None] * (plan.n_act - 1)
    ok(beta=beta)
    bad = dict(good, loc=torch.full((Bd, J), n0k, dtype=torch.int64, device=DEV))
    with pytest.raises(ValueError, match='outside its layer'):
        ok(beta=[bad] + [None] * (plan.n_act - 1))
    with pytest.raises(ValueError, match=r'\[Bd, J\]'):
        ok(beta=[dict(good, sign=torch.ones(Bd, J + 1, device=DEV))] + [None] * (plan.n_act - 1))
    with pytest.raises(TypeError, match='CUDA'):
        ok(beta=[dict(good, loc=good['loc'].cpu())] + [None] * (plan.n_act - 1))
    assert capi.launch_count() == n0
