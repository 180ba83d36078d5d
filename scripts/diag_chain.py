"""Diagnostic (not a test): per-CTA timeline of the chain pass kernel (clock64 stamps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from neuralsat_b200 import capi, synth
from neuralsat_b200.graph import nodes_to, trace_module

dev = 'cuda'
Bd = int(sys.argv[1]) if len(sys.argv) > 1 else 512
net = synth.build_network('mnistfc_256x4')
nodes = trace_module(net, (1, 1, 28, 28))
plan = capi.Plan(nodes_to(nodes, dev))
b = synth.make_batch(nodes, Bd, 0.02, 0, dev)
L = capi.lib()
use_beta = len(sys.argv) > 2 and sys.argv[2] == 'beta'
fn = lambda: plan.crown_pass(b['C'], b['x_L'], b['x_U'], b['lower'], b['upper'], b['alpha'], None, b['beta'] if use_beta else None, want_lA=True)
buf = torch.zeros(64 * 4096, dtype=torch.int64, device=dev)
fn(); torch.cuda.synchronize()
L.cb_debug_tc_times(buf.data_ptr())
fn(); torch.cuda.synchronize()
L.cb_debug_tc_times(None)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    fn()
e1.record(); torch.cuda.synchronize()
print(f'pass (F1, lA out) {e0.elapsed_time(e1) * 100:.1f} us per call, Bd {Bd}')
t = buf.view(-1, 64).cpu()
t = t[t[:, 0] != 0]
print("CTAs", t.shape[0])
fin = (t[:, 62] - t[:, 0]).float()
st = (t[:, 0] - t[:, 0].min()).float()
print(f'MMA warp finish (cycles after the CTA\'s own start): min {fin.min():.0f} mean {fin.mean():.0f} max {fin.max():.0f}; '
      f'CTA start spread: max {st.max():.0f} cycles; sorted finish deciles: '
      + ' '.join(f'{v:.0f}' for v in fin.sort().values[::max(1, t.shape[0] // 10)]))
for cta in (0, t.shape[0] // 2):
    r = t[cta]
    t0 = int(r[0])
    print(f'CTA {cta}')
    print(f'  MMA warp finished at {int(r[62]) - t0}')
    for jc in range(10):
        m = [int(r[1 + 3 * jc + i]) - t0 if r[1 + 3 * jc + i] else None for i in range(3)]
        e = [int(r[32 + 2 * jc + i]) - t0 if r[32 + 2 * jc + i] else None for i in range(2)]
        if m[0] is None and e[0] is None:
            continue
        print(f'  job {jc}: mma start {m[0]}  x ready {m[1]}  mma issued {m[2]} | epi acc_full {e[0]}  epi done {e[1]}')
