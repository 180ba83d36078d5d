"""Diagnostic (not a test): divergence of TC vs SIMT trajectories with the iteration count."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from fixtures import keyed_inputs, load_fixture
from neuralsat_b200 import capi
from neuralsat_b200.graph import activation_indices, nodes_to, preact_indices
from test_cuda_parity import _to_lists

name = 'mnist_fc'
fx, model, nodes = load_fixture(name)
os.environ['CROWN_B200_DISABLE_TC'] = '0'
plan_tc = capi.Plan(nodes_to(nodes, 'cuda'))
os.environ['CROWN_B200_DISABLE_TC'] = '1'
plan_sm = capi.Plan(nodes_to(nodes, 'cuda'))
ent = fx['f2'][0]
k = keyed_inputs(nodes, ent)
# single gradient comparison first
lower, upper, alpha, pos, beta = _to_lists(nodes, k)
r1 = plan_tc.crown_grad(k['C'].cuda(), k['x_L'].cuda(), k['x_U'].cuda(), lower, upper, alpha, pos, beta)
r2 = plan_sm.crown_grad(k['C'].cuda(), k['x_L'].cuda(), k['x_U'].cuda(), lower, upper, alpha, pos, beta)
print('lb diff', (r1[0]-r2[0]).abs().max().item(), 'lb', r2[0].flatten()[:4].tolist())
for j in range(len(r1[2])):
    ga1, ga2 = r1[2][j], r2[2][j]
    d = (ga1-ga2).abs()
    print(f' grad_alpha {j}: maxabs {ga2.abs().max():.3e} maxdiff {d.max():.3e} rel-to-max {d.max()/ga2.abs().max():.2e}; '
          f'nonzero mismatch {(int(((ga1==0)!=(ga2==0)).sum()))}; lA diff {(r1[1][j]-r2[1][j]).abs().max():.3e} lA max {r2[1][j].abs().max():.3e}')
    if r1[3][j] is not None:
        print(f' grad_beta {j}: maxdiff {(r1[3][j]-r2[3][j]).abs().max():.3e} max {r2[3][j].abs().max():.3e}')
for it in [1, 2, 3, 4, 6, 10, 20]:
    res = {}
    for tag, plan in (('tc', plan_tc), ('simt', plan_sm)):
        lower, upper, alpha, pos, beta = _to_lists(nodes, k)
        lb, lA, n_iter = plan.optimize(k['C'].cuda(), k['x_L'].cuda(), k['x_U'].cuda(), lower, upper, alpha, pos, beta,
                                       k['rhs'].cuda(), iteration=it, early_stop=False, early_stop_patience=1000)
        res[tag] = (lb.cpu(), [a.cpu() for a in alpha])
    dl = (res['tc'][0] - res['simt'][0]).abs().max().item()
    da = max((a - b).abs().max().item() for a, b in zip(res['tc'][1], res['simt'][1]))
    print(f'iteration={it}: lb diff {dl:.3e}  alpha maxdiff {da:.3e}')
