"""Where one DeviceBaB.step goes: per-phase wall clock (synchronised) and the library's per-kernel-class CUDA-event times.
usage: prof_step.py WORKLOAD BD"""
import os, sys, time
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


class A:
    no_profile = False


res = {}
import types
args = types.SimpleNamespace()
# reuse the bench's construction of the synthetic store, instrumented
from neuralsat_b200 import domain_store as ds
orig_step = ds.DeviceBaB.step


def timed_step(self, batch, decisions=None):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = orig_step(self, batch, decisions)
    torch.cuda.synchronize(); res.setdefault('step', []).append(time.perf_counter() - t0)
    return out


ds.DeviceBaB.step = timed_step
capi.profile_enable(os.environ.get('NOPROF') != '1')       # NOPROF=1: the un-instrumented rate
r = bench.e2e_device_store(args, None, 0, 1, w, nodes, plan, batch, bd, 4, 2, None, {}, False)
capi.profile_enable(False)
prof = capi.profile_collect()
n = len(res['step'])
print('e2e', r['value'], 'ms/step', r['ms_per_step'], 'steps timed', n, 'mean sync step ms', 1e3 * sum(res['step'][2:]) / max(1, n - 2))
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
    print(f"  {k:14s} {v['ms'] / n:8.2f} ms/step  {v['launches'] // n} launches")
