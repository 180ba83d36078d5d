"""Helpers to turn golden fixtures (tests/golden/*.pt, made by oracle/gen_golden.py from the
unmodified reference) into the keyed inputs of the oracle / the CUDA path."""
import os

import torch

from models import build_model
from neuralsat_b200.graph import trace_module, activation_indices, preact_indices

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def load_fixture(name):
    fx = torch.load(os.path.join(GOLDEN, f'{name}.pt'), weights_only=False)
    model, in_shape = build_model(fx['model'])
    model.load_state_dict(fx['state_dict'])
    model.eval()
    nodes = trace_module(model, (1, *in_shape))
    return fx, model, nodes


def keyed_inputs(nodes, ent):
    """record -> dict(C, x_L, x_U, lower, upper, alpha, alpha_index, beta, rhs)."""
    acts = activation_indices(nodes)
    pres = preact_indices(nodes)
    d = {
        'C': ent['C'], 'x_L': ent['x_L'], 'x_U': ent['x_U'],
        'lower': {pres[k]: ent['lower'][k] for k in range(len(pres))},
        'upper': {pres[k]: ent['upper'][k] for k in range(len(pres))},
        'alpha': {acts[k]: ent['alpha'][k] for k in range(len(acts))},
        'alpha_index': {acts[k]: ent['alpha_index'][k] for k in range(len(acts))},
        'beta': None, 'rhs': ent.get('rhs'),
    }
    if ent.get('enable_beta') and 'beta' in ent:
        d['beta'] = {pres[k]: ent['beta'][k] for k in range(len(pres))}
    return d
