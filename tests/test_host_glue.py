"""CPU: host-side glue of the facade (no kernels): split application, history doubling, SparseBeta
layout, nested option update, graph tracing order — checked against the abstractor-level golden
records of the unmodified reference (tests/golden/*_abs.pt, oracle/gen_golden.py)."""
import os

import pytest
import torch

from fixtures import GOLDEN
from models import build_model
from neuralsat_b200 import abstractor as ab_mod
from neuralsat_b200.bounded_module import BoundedModule, SparseBeta, stop_criterion_batch_any, _threshold_of
from neuralsat_b200.graph import activation_indices, preact_indices, trace_module


def _bare_abstractor(device='cpu'):
    ab = ab_mod.NetworkAbstractor.__new__(ab_mod.NetworkAbstractor)
    ab.device = device
    ab.input_split = False
    return ab


@pytest.mark.parametrize('name', ['fc_small', 'conv_small', 'fc_sigmoid'])
def test_split_application_and_histories_match_reference(name):
    fx = torch.load(os.path.join(GOLDEN, f'{name}_abs.pt'), weights_only=False)
    ab = _bare_abstractor()
    for rec in fx['records']:
        p, out, dec = rec['params'], rec['out'], rec['decisions']
        B = len(dec)
        new = ab.hidden_split_idx(p['lower_bounds'], p['upper_bounds'], [list(d) for d in dec])
        for k in p['lower_bounds']:
            if k == 'final':
                continue
            assert torch.equal(new[k][0], out['lower_bounds'][k]), k
            assert torch.equal(new[k][1], out['upper_bounds'][k]), k
        hist = ab.update_histories(p['histories'], [list(d) for d in dec])
        assert len(hist) == 2 * B
        for h, h_ref in zip(hist, out['histories']):
            for k in h_ref:
                for a, b in zip(h[k], h_ref[k]):
                    assert torch.equal(torch.as_tensor(a).float(), torch.as_tensor(b).float())


def test_sparse_beta_layout():
    hist = [{'pre0': (torch.tensor([3, 5]), torch.tensor([1., -1.]), torch.tensor([0., 0.]))},
            {'pre0': (torch.tensor([7]), torch.tensor([-1.]), torch.tensor([0.]))}]
    sb = SparseBeta((2, 2), bias=False, betas=[torch.tensor([0.5]), None], device='cpu')
    sb.apply_splits(hist, 'pre0')
    assert sb.val.tolist() == [[0.5, 0.0], [0.0, 0.0]]
    assert sb.loc.tolist() == [[3, 5], [7, 0]]
    assert sb.sign.tolist() == [[1.0, -1.0], [-1.0, 0.0]]     # padded entries have sign 0
    assert sb.bias is None
    sb2 = SparseBeta((2, 2), bias=True, device='cpu')
    sb2.apply_splits(hist, 'pre0')
    assert sb2.bias.shape == (2, 2)


def test_stop_criterion_threshold_roundtrip():
    rhs = torch.tensor([[0.0], [1.0]])
    f = stop_criterion_batch_any(rhs)
    assert f(torch.tensor([[0.5], [0.5]])).flatten().tolist() == [True, False]
    assert _threshold_of(f) is rhs
    ref_style = (lambda thr: (lambda x: (x > thr).any(dim=1, keepdim=True)))(rhs)   # AL/utils.py:87-93
    assert _threshold_of(ref_style) is rhs


def test_set_bound_opts_is_a_nested_update():
    bm = BoundedModule.__new__(BoundedModule)
    torch.nn.Module.__init__(bm)
    bm.bound_opts = {'optimize_bound_args': {'iteration': 20, 'lr_alpha': 0.5}, 'crown_batch_size': 1}
    bm.set_bound_opts({'optimize_bound_args': {'iteration': 5}, 'crown_batch_size': 7})
    assert bm.bound_opts == {'optimize_bound_args': {'iteration': 5, 'lr_alpha': 0.5}, 'crown_batch_size': 7}


def test_trace_order_is_program_order():
    model, in_shape = build_model('resnet_bn_small')
    nodes = trace_module(model, (1, *in_shape))
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    assert len(acts) == 4 and all(nodes[a]['op'] == 'relu' for a in acts)
    assert acts == sorted(acts) and [nodes[a]['in'][0] for a in acts] == pres
    assert any(nd['op'] == 'add' for nd in nodes) and any(nd['op'] == 'batchnorm2d' for nd in nodes)


def test_facade_needs_cuda():
    model, in_shape = build_model('fc_small')
    with pytest.raises(RuntimeError, match='no CPU path'):
        BoundedModule(model, torch.zeros(1, *in_shape), device='cpu')


@pytest.mark.parametrize('op', ['sigmoid', 'tanh'])
def test_sshape_tables_match_oracle(op):
    """The product's tangent tables (uploaded to the GPU as plan constants) are bit-identical to the
    oracle's restatement of AL/operators/tanh.py:65-130, and so are the looked-up initial points."""
    from neuralsat_b200 import sshape_tables as st
    from oracle import sshape_oracle as sso
    dl, du = st.tangent_tables(op, 'cpu')
    rl, ru = sso.tables(op)
    assert dl.shape == (50005,) and torch.equal(dl, rl) and torch.equal(du, ru)
    g = torch.Generator().manual_seed(0)
    l = torch.randn(64, 7, generator=g) * 3 - 1
    u = l + torch.rand(64, 7, generator=g) * 4
    a, b = st.lookup_points(op, l, u)
    ra, rb = sso.lookup(op, l, u)
    assert torch.equal(a, ra) and torch.equal(b, rb)


def test_ragged_fill_equals_the_per_row_loop():
    """`_ragged_fill` (one concatenation + one indexed store) against the per-domain loop of the reference
    (AL/beta_crown.py:25-42) on mixed inputs: tensors, lists, empty rows, None."""
    from neuralsat_b200.bounded_module import _ragged_fill
    g = torch.Generator().manual_seed(3)
    rows = []
    for i in range(37):
        n = int(torch.randint(0, 6, (1,), generator=g))
        kind = i % 4
        vals = torch.randn(n, generator=g)
        rows.append(None if kind == 0 and n == 0 else (vals.tolist() if kind == 1 else vals))
    rows[5] = None
    rows[6] = torch.empty(0)
    dst = torch.full((37, 6), -7.0)
    ref = dst.clone()
    for i, r in enumerate(rows):
        if r is not None and len(r):
            ref[i, :len(r)] = torch.as_tensor(r)
    _ragged_fill(dst, rows)
    assert torch.equal(dst, ref)
    # integer destination (the `loc` table), values converted to its dtype
    long_rows = [None if r is None else torch.as_tensor(r).mul(10).long() for r in rows]
    long_dst, long_ref = torch.zeros(37, 6, dtype=torch.long), torch.zeros(37, 6, dtype=torch.long)
    for i, r in enumerate(long_rows):
        if r is not None and len(r):
            long_ref[i, :len(r)] = r
    _ragged_fill(long_dst, long_rows)
    assert torch.equal(long_dst, long_ref)
    _ragged_fill(dst, [None] * 37)                      # nothing to do: untouched
    assert torch.equal(dst, ref)


def test_get_beta_returns_each_domains_prefix():
    """`get_beta` (abstractor/utils.py:119-131): per domain and layer the first n_splits values of the padded table."""
    from types import SimpleNamespace
    ab = _bare_abstractor()
    val = {'a': torch.arange(12.).reshape(3, 4), 'b': torch.arange(6.).reshape(3, 2) + 100}
    ab.net = {k: SimpleNamespace(sparse_betas=[SimpleNamespace(val=v)]) for k, v in val.items()}
    n_splits = [{'a': 2, 'b': 0}, {'a': 0, 'b': 2}, {'a': 4, 'b': 1}]
    out = ab.get_beta(n_splits)
    assert len(out) == 3
    for i, ns in enumerate(n_splits):
        for k, n in ns.items():
            assert torch.equal(out[i][k], val[k][i, :n]), (i, k)
    assert ab.get_beta([]) == []
