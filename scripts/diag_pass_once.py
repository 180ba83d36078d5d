"""Diagnostic (not a test): a few F1 passes of the MNIST-FC workload, for ncu captures."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to, trace_module
dev = 'cuda'
Bd = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
net = synth.build_network('mnistfc_256x4')
nodes = trace_module(net, (1, 1, 28, 28))
plan = capi.Plan(nodes_to(nodes, dev))
b = synth.make_batch(nodes, Bd, 0.02, 0, dev)
for _ in range(4):
    plan.crown_pass(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, None, want_lA=True)
torch.cuda.synchronize()
