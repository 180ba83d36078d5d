"""Diagnostic (not a test): TC path vs SIMT path vs reference fixture on F2."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
from fixtures import keyed_inputs, load_fixture
from neuralsat_b200 import capi
from neuralsat_b200.graph import activation_indices, nodes_to, preact_indices
from test_cuda_parity import _to_lists

for name in ['fc_small', 'mnist_fc']:
    fx, model, nodes = load_fixture(name)
    os.environ['CROWN_B200_DISABLE_TC'] = '0'
    plan_tc = capi.Plan(nodes_to(nodes, 'cuda'))
    os.environ['CROWN_B200_DISABLE_TC'] = '1'
    plan_sm = capi.Plan(nodes_to(nodes, 'cuda'))
    os.environ['CROWN_B200_DISABLE_TC'] = '0'
    print(name, 'tc contractions', plan_tc.tc_contractions, plan_sm.tc_contractions)
    for ei, ent in enumerate(fx['f2']):
        k = keyed_inputs(nodes, ent)
        outs = {}
        for tag, plan in (('tc', plan_tc), ('simt', plan_sm)):
            lower, upper, alpha, pos, beta = _to_lists(nodes, k)
            lb, lA, n_iter = plan.optimize(k['C'].cuda(), k['x_L'].cuda(), k['x_U'].cuda(), lower, upper, alpha, pos, beta,
                                           k['rhs'].cuda(), iteration=ent['iteration'], lr_alpha=ent['lr_alpha'],
                                           lr_beta=ent['lr_beta'], lr_decay=ent['lr_decay'], enable_beta=ent['enable_beta'],
                                           early_stop=True)
            outs[tag] = (lb.cpu(), [a.cpu() for a in alpha], [t.cpu() for t in lA], [b['val'].cpu() for b in beta], n_iter)
        ref = ent['out_lb']
        for tag in ('tc', 'simt'):
            lb, alpha, lA, bv, n_iter = outs[tag]
            print(f' rec {ei} {tag}: n_iter {n_iter} lb err {(lb-ref).abs().max():.3e} (|ref| {ref.abs().max():.3f})')
            for j in range(len(alpha)):
                da = (alpha[j] - ent['out_alpha'][j]).abs()
                bad = da > 2e-3 + 1e-3 * ent['out_alpha'][j].abs()
                dl = (lA[j] - ent['out_lA'][j]).abs().max()
                msg = f'   act {j}: alpha maxdiff {da.max():.3e} bad {int(bad.sum())}/{bad.numel()} lA maxdiff {dl:.3e}'
                if bad.any():
                    idx = bad[0].nonzero()[:4]
                    for t in idx:
                        s1, b_, c_ = [int(v) for v in t]
                        ai = k['alpha_index'][activation_indices(nodes)[j]]
                        neuron = int(ai[c_]) if ai is not None else c_
                        p = preact_indices(nodes)[j]
                        msg += (f'\n      b={b_} col={c_} neuron={neuron} ours={alpha[j][0,s1,b_,c_]:.4f} ref={ent["out_alpha"][j][0,s1,b_,c_]:.4f}'
                                f' lA_ref={ent["out_lA"][j].reshape(ent["out_lA"][j].shape[0], ent["out_lA"][j].shape[1], -1)[0,b_,neuron]:.3e}'
                                f' l={k["lower"][p].reshape(k["lower"][p].shape[0],-1)[b_,neuron]:.3e} u={k["upper"][p].reshape(k["upper"][p].shape[0],-1)[b_,neuron]:.3e}')
                print(msg)
            for j in range(len(bv)):
                db = (bv[j] - ent['out_beta_val'][j]).abs().max() if bv[j].numel() else 0.
                print(f'   beta {j}: maxdiff {float(db):.3e}')
