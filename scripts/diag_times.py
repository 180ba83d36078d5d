"""Diagnostic (not a test): per-CTA phase timeline of one tensor-core launch (clock64 stamps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to, trace_module

dev = 'cuda'
net = synth.build_network('mnistfc_256x4')
nodes = trace_module(net, (1, 1, 28, 28))
plan = capi.Plan(nodes_to(nodes, dev))
Bd = 8192
b = synth.make_batch(nodes, Bd, 0.02, 0, dev)
L = capi.lib()
def run(tag, fn, ncta):
    buf = torch.zeros(8 * 4096, dtype=torch.int64, device=dev)
    fn()  # warm
    torch.cuda.synchronize()
    L.cb_debug_tc_times(buf.data_ptr())
    fn()
    torch.cuda.synchronize()
    L.cb_debug_tc_times(None)
    t = buf.view(-1, 8).cpu()
    t = t[t[:, 0] != 0]
    print(tag, 'CTAs recorded (last launch overwrites earlier ones):', t.shape[0])
    t0 = t[:, 0:1]
    d = (t[:, :8] - t0).float()
    names = ['start', 'setup', 'stage0', 'mma_issued', 'ew_landed', 'acc_ready', 'epi_math', 'epi_out']
    for i, n in enumerate(names):
        print(f'   {n:11s} mean {d[:, i].mean():9.0f}  min {d[:, i].min():9.0f}  max {d[:, i].max():9.0f} cycles')
    span = 0
    print('   whole-launch span', span, 'cycles; start spread', (t[:, 0].max() - t[:, 0].min()).item())
# F1 pass: the LAST tc launch is the concretize (N=784); run a grad to see GRAD (last = layer 4 grad)
run('pass (last launch = CONCRETIZE 256->784)', lambda: plan.crown_pass(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, None, want_lA=False), 0)
run('grad (last launch = GRAD 256->256)', lambda: plan.crown_grad(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta']), 0)
