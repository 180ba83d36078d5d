"""Diagnostic (not a test): chain pass timing under CROWN_B200_EXP experiment flags (results invalid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to, trace_module
dev = 'cuda'
Bd = 8192
net = synth.build_network('mnistfc_256x4')
nodes = trace_module(net, (1, 1, 28, 28))
plan = capi.Plan(nodes_to(nodes, dev))
b = synth.make_batch(nodes, Bd, 0.02, 0, dev)
for want_lA in (True, False):
    fn = lambda: plan.crown_pass(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'], want_lA=want_lA)
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print('EXP', os.environ.get('CROWN_B200_EXP', '0'), 'want_lA', want_lA, 'us per pass', round(e0.elapsed_time(e1) / 20 * 1e3, 1))
