"""Device-resident domain store + branching + BaB step against the unmodified reference:

* split decisions: `DeviceBaB.branch` (BaBSR scores, top-k, batched look-ahead, arg-max) must pick exactly the
  decisions `DecisionHeuristic.smart_hidden_branching` picked on the same picked domains (tests/golden/*_abs.pt);
* one iteration: children, pruning and the records appended to the store against `NetworkAbstractor.forward` +
  `DomainsList.add` of the reference;
* whole runs: the reference's BaB loop recorded to the end (oracle/gen_root_golden.py:bab_run) - same verdict, same
  queue length after every iteration, same decisions."""
import os

import pytest
import torch

from fixtures import GOLDEN
from models import build_model

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _net(fx):
    from neuralsat_b200.bounded_module import BoundedModule
    model, in_shape = build_model(fx['model'])
    model.load_state_dict(fx['state_dict'])
    net = BoundedModule(model.eval(), torch.zeros(1, *in_shape), device=DEV)
    return net


def _names(net):
    m = {'final': net.final_name}
    for k, n in enumerate(net.split_nodes):
        m[f'pre{k}'] = n.name
    for k, a in enumerate(net.perturbed_optimizable_activations):
        m[f'act{k}'] = a.name
    return m


def _results(net, d, alpha_index):
    """canonical dict (oracle/gen_golden.py:_canon_results) -> AbstractResults with this facade's node names."""
    from neuralsat_b200.abstractor import AbstractResults
    nm = _names(net)
    for a, idx in zip(net.perturbed_optimizable_activations, alpha_index):
        a.alpha_indices = None if idx is None else (idx.to(DEV),)
        a._alpha_pos = None
    ren = lambda v: None if v is None else {nm[k]: x for k, x in v.items()}
    hist = None if d.get('histories') is None else [{nm[k]: v for k, v in h.items()} for h in d['histories']]
    betas = None if d.get('betas') is None else [None if b is None else {nm[k]: v for k, v in b.items()} for b in d['betas']]
    slopes = {nm[k]: {nm[kk]: x for kk, x in v.items()} for k, v in d['slopes'].items()}
    return AbstractResults(objective_ids=d['objective_ids'], output_lbs=d['output_lbs'], lAs=ren(d['lAs']),
                           lower_bounds=ren(d['lower_bounds']), upper_bounds=ren(d['upper_bounds']),
                           input_lowers=d['input_lowers'], input_uppers=d['input_uppers'], slopes=slopes, betas=betas,
                           histories=hist, cs=d['cs'], rhs=d['rhs'])


@pytest.mark.parametrize('name,topk', [('fc_small', 2), ('conv_small', 2)])
def test_decisions_and_step_match_reference(name, topk):
    from neuralsat_b200.domain_store import DeviceBaB, DeviceDomainStore
    fx = torch.load(os.path.join(GOLDEN, f'{name}_abs.pt'), weights_only=False)
    net = _net(fx)
    for rec in fx['records']:
        params = _results(net, rec['params'], rec['alpha_index'])
        B = len(rec['decisions'])
        store = DeviceDomainStore(net, params)
        assert len(store) == B
        bab = DeviceBaB(net, store, decision_topk=topk)
        pick = store.pick_out(B)
        layer, neuron = bab.branch(pick)
        got = [(f'pre{int(l)}', int(n)) for l, n in zip(layer.cpu(), neuron.cpu())]
        ref = [(d[0], d[1]) for d in rec['decisions']]
        assert got == ref, (got, ref)
        # the full step on the same picked set (decisions recomputed by the heuristic)
        store.n += B                     # put the picked records back: step() picks them itself
        store.visited -= B
        info = bab.step(B)
        out = rec['out']
        lb_ref = out['output_lbs']
        lb = bab.last['lb'].cpu()
        assert torch.allclose(lb, lb_ref, rtol=1e-5, atol=1e-5 * max(1.0, float(lb_ref.abs().max()))), (lb - lb_ref).abs().max()
        keep_ref = (lb_ref <= out['rhs']).all(1)
        assert info['kept'] == int(keep_ref.sum()) and len(store) == info['kept']
        # appended records == the reference's surviving children, in the same order (domains_list.py:240-300)
        idx = keep_ref.nonzero().flatten()
        n = len(store)
        assert torch.allclose(store.lb[:n].cpu(), lb_ref[idx], rtol=1e-5, atol=1e-5)
        for k, p in enumerate(net.split_nodes):
            assert torch.equal(store.lower[k][:n].cpu().view(n, -1), out['lower_bounds'][f'pre{k}'][idx].view(n, -1))
            assert torch.equal(store.upper[k][:n].cpu().view(n, -1), out['upper_bounds'][f'pre{k}'][idx].view(n, -1))
            for row, j in enumerate(idx.tolist()):
                loc, sign, _ = out['histories'][j][f'pre{k}']
                c = int(store.h_cnt[k][row])
                assert c == len(loc)
                assert torch.equal(store.h_loc[k][row, :c].cpu().long(), torch.as_tensor(loc).long())
                assert torch.equal(store.h_sign[k][row, :c].cpu(), torch.as_tensor(sign).float())
        for k, a in enumerate(net.perturbed_optimizable_activations):
            ref_a = out['slopes'][f'act{k}']['final'][0, 0][idx].reshape(n, -1)          # fp16, as the reference stores them
            got_a = store.alpha[k][:n].cpu()
            bad = ((got_a.float() - ref_a.float()).abs() > 2e-3).float().mean()
            assert bad <= 0.02, bad


@pytest.mark.parametrize('name', ['fc_small', 'conv_small'])
def test_whole_run_matches_reference(name):
    """Identical verdict, queue lengths and split decisions over a whole BaB run started from the reference's root."""
    from neuralsat_b200.domain_store import DeviceBaB, DeviceDomainStore
    fx = torch.load(os.path.join(GOLDEN, f'bab_{name}.pt'), weights_only=False)
    net = _net(fx)
    root = _results(net, fx['root'], fx['alpha_index'])
    store = DeviceDomainStore(net, root)
    bab = DeviceBaB(net, store, decision_topk=fx['topk'])
    same_decisions = 0
    total = 0
    queue = []
    for it in fx['iterations']:
        assert len(store) > 0
        info = bab.step(fx['batch'])
        queue.append(info['remaining'])
        got = [(f'pre{int(l)}', int(n)) for l, n in zip(bab.last['layer'].cpu(), bab.last['neuron'].cpu())]
        total += len(got)
        same_decisions += sum(1 for a, b in zip(got, it['decisions']) if a == b) if len(got) == len(it['decisions']) else 0
    assert queue == [it['remaining'] for it in fx['iterations']], (queue, [it['remaining'] for it in fx['iterations']])
    assert same_decisions == total, (same_decisions, total)
    verdict = 'unsat' if len(store) == 0 else 'unknown'
    assert verdict == fx['verdict'] and store.visited == fx['visited']


def test_facade_started_run_verifies():
    """No reference state at all: ONNX-free model -> initialize -> device BaB loop -> 'unsat', and the proof is sound on
    samples (every sampled output satisfies the property)."""
    from types import SimpleNamespace
    from neuralsat_b200.abstractor import NetworkAbstractor
    from neuralsat_b200.domain_store import DeviceBaB, DeviceDomainStore
    model, in_shape = build_model('fc_small')
    torch.manual_seed(3)
    for p in model.parameters():
        torch.nn.init.normal_(p, std=0.3)
    model.eval()
    g = torch.Generator().manual_seed(0)
    x0 = torch.rand(1, *in_shape, generator=g)
    with torch.no_grad():
        y = model(x0)
    label = int(y.argmax())
    others = [j for j in range(y.shape[1]) if j != label]
    eps = 0.04
    C = torch.zeros(len(others), 1, y.shape[1])
    for r, j in enumerate(others):
        C[r, 0, label], C[r, 0, j] = 1., -1.
    N = len(others)
    obj = SimpleNamespace(lower_bounds=(x0 - eps).flatten(1).repeat(N, 1), upper_bounds=(x0 + eps).flatten(1).repeat(N, 1),
                          cs=C, rhs=torch.zeros(N, 1), ids=torch.arange(N) + 3)
    ab = NetworkAbstractor(model, (1, *in_shape), 'crown-optimized', input_split=False, device=DEV)
    ab.setup(obj)
    root = ab.initialize(obj)
    with torch.no_grad():
        xs = x0 + (torch.rand(4096, *in_shape, generator=g) * 2 - 1) * eps
        margins = torch.einsum('rsn,kn->krs', C, model(xs)).min()
    if root.lower_bounds is None:
        assert margins > 0
        return
    store = DeviceDomainStore(ab.net, root)
    bab = DeviceBaB(ab.net, store, decision_topk=3)
    verdict = bab.run(batch=64, max_iterations=200)
    if verdict == 'unsat':
        assert margins > 0, 'proved a property that a sample violates'
    assert verdict in ('unsat', 'unknown')
