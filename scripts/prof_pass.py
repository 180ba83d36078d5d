"""Profiling helper: a few pass + gradient launches of one workload (for ncu launch lists).
usage: prof_pass.py WORKLOAD BD [REPS]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to, trace_module

w, bd = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
wl = synth.WORKLOADS[w]
nodes = synth.build_nodes(w, 0)
plan = capi.Plan(nodes_to(nodes, 'cuda'))
b = synth.make_batch(nodes, bd, wl['eps'], 0, 'cuda', bounds=wl.get('bounds', 'ibp'))
for _ in range(reps):
    plan.crown_grad(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'])
torch.cuda.synchronize()
print('done')
