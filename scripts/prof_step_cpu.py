"""cProfile of the host side of DeviceBaB.step (bench.e2e_device_store).  usage: prof_step_cpu.py WORKLOAD BD"""
import cProfile, os, pstats, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to

w, bd = sys.argv[1], int(sys.argv[2])
wl = synth.WORKLOADS[w]
nodes = synth.build_nodes(w, 0)
plan = capi.Plan(nodes_to(nodes, 'cuda'))
batch = synth.make_batch(nodes, bd, wl['eps'], 0, 'cuda', bounds=wl.get('bounds', 'ibp'))
args = types.SimpleNamespace()
bench.e2e_device_store(args, None, 0, 1, w, nodes, plan, batch, bd, 2, 2, None, {}, False)
pr = cProfile.Profile()
pr.enable()
r = bench.e2e_device_store(args, None, 0, 1, w, nodes, plan, batch, bd, 20, 2, None, {}, False)
pr.disable()
print('e2e', r['value'], r['ms_per_step'])
pstats.Stats(pr).sort_stats('tottime').print_stats(18)
