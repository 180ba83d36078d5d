"""The UNMODIFIED reference's own path timed on the host cores: `NetworkAbstractor.forward(decisions, domain_params)`
(NS/abstractor/abstractor.py:403-406 -> _forward_hidden :244-344 -> auto_LiRPA compute_bounds('crown-optimized')).

TEST / BENCH INFRASTRUCTURE ONLY (bench.py --impl reference and its cpu_baseline leg).  The reference tree is found at
$NEURALSAT_REFERENCE, else baseline/_ref/neuralsat-pt201 (staged by scripts/stage_reference.sh, git-ignored, travels
to the GPU box), else /root/reference/neuralsat-pt201.

The workload is the bench's: a synthetic batch of sub-domains of one architecture (neuralsat_b200.synth).  The
reference module gets its slopes the way `set_slope` installs them in the BaB loop (one [2,1,B,n] tensor per ReLU for
the output start node, dense), the parents' histories / betas as the per-domain dicts of its API, and one new split
decision per parent; early stop is defeated by an infinite threshold so that every step does the full 20 iterations.
"""
import os
import sys
import time
from collections import OrderedDict

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def locate():
    for p in (os.environ.get('NEURALSAT_REFERENCE'), os.path.join(ROOT, 'baseline', '_ref', 'neuralsat-pt201'),
              '/root/reference/neuralsat-pt201'):
        if p and os.path.isdir(os.path.join(p, 'auto_LiRPA')):
            return p
    return None


class ReferenceArm:
    def __init__(self, workload: str, n_parents: int, seed: int = 0):
        ref = locate()
        if ref is None:
            raise RuntimeError('reference tree not staged (scripts/stage_reference.sh)')
        os.environ['NEURALSAT_REFERENCE'] = ref
        if ROOT not in sys.path:
            sys.path.insert(0, ROOT)
        from oracle import ref_bootstrap as rb
        rb.REF_ROOT = ref
        rb.bootstrap()
        from abstractor.abstractor import NetworkAbstractor
        from onnx2pytorch.convert.model import ConvertModel
        from util.misc.result import AbstractResults
        from setting import Settings
        from neuralsat_b200 import synth
        from neuralsat_b200.graph import activation_indices, preact_indices
        Settings.use_restart = False
        self.AbstractResults = AbstractResults
        wl = synth.WORKLOADS[workload]
        model = synth.build_network(workload, seed=0)
        nodes = synth.build_nodes(workload, seed=0, fold_bn=False)
        B = n_parents
        b = synth.make_batch(nodes, B, wl['eps'], seed=seed, device='cpu', bounds=wl.get('bounds', 'ibp'))
        in_shape = (1, *wl['in_shape'])
        obj = rb.Objective(b['x_L'][:1].flatten(1), b['x_U'][:1].flatten(1), b['C'][:1], torch.zeros(1, 1), torch.arange(1))
        ab = NetworkAbstractor(ConvertModel(model).eval(), in_shape, 'crown-optimized', input_split=False, device='cpu')
        ab._init_module(mode='matrix', objective=obj)
        ab.mode, ab.method = 'matrix', 'crown-optimized'
        net = ab.net
        x = ab.new_input(x_L=b['x_L'][:1], x_U=b['x_U'][:1])
        net(x)
        net.get_split_nodes(input_split=False)
        # one plain CROWN pass so that every node knows its shapes / perturbation flags, then the split-node list
        with torch.no_grad():
            net.compute_bounds(x=(x,), C=b['C'][:1], method='backward')
        net.get_split_nodes(input_split=False)
        # initialize() leaves init_alpha=False in auto_LiRPA's option dict for every later beta step
        # (NS/abstractor/params.py:39, SURVEY.md section 5 "Config / flags"); without the root call it has to be set here
        net.set_bound_opts({'optimize_bound_args': {'init_alpha': False, 'early_stop_patience': 10 ** 6}})    # constant work: all 20 iterations
        acts = list(net.perturbed_optimizable_activations)
        pres = list(net.split_nodes)
        assert len(acts) == len(activation_indices(nodes)) == len(pres)
        final = net.final_name
        # dense output-node slopes, as set_slope would install them (NS/abstractor/utils.py:63-74)
        for m, a in zip(acts, b['alpha']):
            m.alpha = OrderedDict({final: a[:, :, :1].clone().requires_grad_()})
            m.alpha_lookup_idx = OrderedDict({final: None})
            m.alpha_indices = None
            m.alpha_size = 2
        self.ab, self.net, self.B = ab, net, B
        slopes = {m.name: {final: a.half()} for m, a in zip(acts, b['alpha'])}
        lower = {p.name: l for p, l in zip(pres, b['lower'])}
        upper = {p.name: u for p, u in zip(pres, b['upper'])}
        hist, betas = [], []
        for i in range(B):
            h, bt = {}, {}
            for p, rec in zip(pres, b['beta']):
                live = rec['sign'][i] != 0
                h[p.name] = (rec['loc'][i][live].clone(), rec['sign'][i][live].clone(), torch.zeros(int(live.sum())))
                bt[p.name] = rec['val'][i][live].clone()
            hist.append(h)
            betas.append(bt)
        g = torch.Generator().manual_seed(seed + 1)
        self.decisions = []
        for i in range(B):
            k = int(torch.randint(0, len(pres), (1,), generator=g))
            n = int(torch.randint(0, lower[pres[k].name][0].numel(), (1,), generator=g))
            self.decisions.append([pres[k].name, n, 0.0])
        self.params = AbstractResults(objective_ids=torch.arange(B), output_lbs=torch.zeros(B, 1),
                                      input_lowers=b['x_L'], input_uppers=b['x_U'], lower_bounds=lower, upper_bounds=upper,
                                      lAs=None, slopes=slopes, betas=betas, histories=hist, cs=b['C'],
                                      rhs=torch.full((B, 1), float('inf')), sat_solvers=None)

    def step(self):
        """One BaB abstraction step of the reference: 2 * B children bounded (20 alpha/beta iterations)."""
        out = self.ab.forward(self.decisions, self.params)
        return out

    def time(self, steps: int, warmup: int = 1) -> float:
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / steps
