#!/usr/bin/env python
"""bench.py — BaB sub-domains bounded per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one alpha/beta-CROWN bounding (form F2: 20 optimiser iterations, early stop disabled so the work is
constant) of one batch of Bd synthetic sub-domains per GPU through the C-ABI (libcrown_b200.so).
The headline workload is the largest single-GPU configuration of BASELINE.json: configs[3], the CIFAR-10 ResNet
`sri_resnet_a` (residual Add, 1x1 stride-2 shortcuts); the other configurations are measured in the same run with
fewer steps and reported under `workloads` (`--workload NAME` makes any of them the headline, `--no-extra` skips them).

  value      sub-domains bounded / s, inputs resident in HBM (capi.Plan.optimize on prepared batches)
  e2e        the same through the public BaB-step API `neuralsat_b200.domain_store.DeviceBaB.step`: the domains live
             in the device-resident store, the host hands over the split decisions of the step from pinned memory
             (H2D) and reads the children's lower bounds and the survivor count back (D2H); children are built,
             bounded, pruned and appended on the device
  e2e_host_buffers   every input from pinned HOST buffers (neuralsat_b200.pipeline.HostPipeline, the round-1 e2e)
  f1         one CROWN pass per sub-domain (form F1);   branching: BaBSR + top-k look-ahead decisions / s
  roofline   dominant kernel class, CUDA-event time per launch measured inside the timed region
  cpu_baseline   the UNMODIFIED reference (NetworkAbstractor.forward on the host cores, kind "reference") when it has
             been staged into baseline/_ref (scripts/stage_reference.sh), else its port oracle/crown_oracle.py
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'bab_subdomains_bounded_per_sec'
UNIT = 'subdomains/s'
ITERATION = 20
HEADLINE = 'sri_resnet_a'
# sub-domains per GPU and step (2*B children), and children per step of the CPU reference arm
DEFAULT_BD = {'mnistfc_256x4': 9472, 'oval21_base': 16384, 'sri_resnet_a': 16384, 'cifar10_2_255': 4096,
              'cifar100_resnet_medium': 2048, 'tinyimagenet_resnet_medium': 256, 'acasxu': 9472}
CPU_SAMPLE = {'mnistfc_256x4': 2048, 'oval21_base': 512, 'sri_resnet_a': 256, 'cifar10_2_255': 64,
              'cifar100_resnet_medium': 16, 'tinyimagenet_resnet_medium': 8, 'acasxu': 2048}
EXTRA = ['mnistfc_256x4', 'oval21_base', 'cifar10_2_255', 'cifar100_resnet_medium']


# The contract is ONE JSON line on stdout.  Libraries write banners there too (NCCL prints its version on the first
# communicator), so file descriptor 1 is pointed at stderr for the whole run and the line goes to the saved descriptor.
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(text):
    os.write(_STDOUT_FD, (text + '\n').encode())


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default=HEADLINE)
    ap.add_argument('--bd', type=int, default=0, help='sub-domains per GPU per step (2*B children); 0 = the workload default')
    ap.add_argument('--cpu-sample', type=int, default=0, help='children per step of the CPU reference arm; 0 = default')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--no-extra', action='store_true', help='headline workload only')
    ap.add_argument('--only-f2', action='store_true', help='skip the F1 / e2e legs (profiling runs)')
    return ap.parse_args()


def config_of(workload, Bd, world):
    """The `config` object of the JSON line; both arms print the same one."""
    from neuralsat_b200 import synth
    wl = synth.WORKLOADS[workload]
    return {'workload': f'{workload} eps={wl["eps"]:.4g} F2 alpha/beta-CROWN step ({ITERATION} it., early stop off)',
            'subdomains_per_gpu_per_step': Bd, 'spec_rows': 1,
            'l2_policy': 'rotating input batches, working set > L2',
            'parallelism': f'domains sharded x{world}, lb all_gather per step' if world > 1 else 'single GPU'}


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return {'hbm_gbs': d['hbm_gbs'], 'bf16_tflops': d['bf16_tflops'],
                'bf16_tflops_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']), 'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0, 'source': 'fallback'}


class NvmlSampler:
    """SM clock / throttle reasons / power through NVML, polled every ~2 ms from a thread: the timed region of the
    default run is ~60 ms, far too short for `nvidia-smi -lms 100` (ClockSampler below, the fallback) to see it."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.stop_flag = False
        self.th = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = int(vis.split(',')[self.index]) if vis and vis.split(',')[self.index].isdigit() else self.index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv = pynvml
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._poll, daemon=True)
            self.th.start()
            return True
        except Exception:
            self.th = None
            return False

    def _poll(self):
        import time
        nv = self.nv
        while not self.stop_flag:
            try:
                self.samples.append((float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)),
                                     int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)),
                                     nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0))
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self.stop_flag = True
        if self.th is not None:
            self.th.join(timeout=2)
        nv = self.nv
        names = {'hw_slowdown': nv.nvmlClocksEventReasonHwSlowdown, 'hw_thermal_slowdown': nv.nvmlClocksEventReasonHwThermalSlowdown,
                 'sw_thermal_slowdown': nv.nvmlClocksEventReasonSwThermalSlowdown, 'sw_power_cap': nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in names.items() if any(r & bit for _, r, _ in self.samples))
        sm = sorted(c for c, _, _ in self.samples)
        med = sm[len(sm) // 2] if sm else None
        pw = max((p for _, _, p in self.samples), default=None)
        return {'sm_mhz': med, 'sm_min_mhz': sm[0] if sm else None, 'sm_max_mhz': self.mx, 'reasons': reasons,
                'samples': len(sm), 'power_w_max': pw, 'how': 'NVML polled every ~2 ms over the timed region'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100',
                 '-i', str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(',')]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'], f[4:8]):
                if v == 'Active':
                    reasons.add(name)
        sm.sort()
        # median over samples taken under load (the upper half of the samples)
        load = sm[len(sm) // 2:] if sm else []
        med = load[len(load) // 2] if load else None
        return {'sm_mhz': med, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm)}


def dist_setup(n):
    if n > 1 or int(os.environ.get('WORLD_SIZE', '1')) > 1:
        import torch.distributed as dist
        rank = int(os.environ.get('RANK', '0'))
        local = int(os.environ.get('LOCAL_RANK', '0'))
        world = int(os.environ.get('WORLD_SIZE', '1'))
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
        return dist, rank, local, world
    torch.cuda.set_device(0)
    return None, 0, 0, 1


def sync_all(dist):
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()


def timed(dist, fn, steps):
    """barrier+sync, CUDA events around `steps` calls of fn(i), max over ranks -> seconds."""
    sync_all(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    return float(ms.item()) / 1e3


def clone_batch(b):
    return {'C': b['C'], 'x_L': b['x_L'], 'x_U': b['x_U'], 'lower': b['lower'], 'upper': b['upper'],
            'alpha': [a.clone() for a in b['alpha']],
            'beta': [dict(bt, val=bt['val'].clone()) for bt in b['beta']]}



def kernel_classes(nodes, work, Bd):
    """kernel class -> ('tensor', flop per F2 step) | ('hbm', bytes per F2 step): the algorithmic work of DESIGN.md."""
    lin = sum(nd['weight'].numel() for nd in nodes if nd['op'] == 'linear')
    lin_last = [nd['weight'].numel() for nd in nodes if nd['op'] == 'linear'][-1]
    conv = work['mac'] - lin
    first = next(nd for nd in nodes if nd['op'] in ('linear', 'conv2d'))
    n_first = first['weight'].numel() * (first['shape'][1] * first['shape'][2] if first['op'] == 'conv2d' else 1)
    P, G = ITERATION, ITERATION - 1
    nr = work['n_relu']
    return {
        'chain_pass': ('tensor', 2.0 * Bd * lin * P), 'chain_grad': ('tensor', 2.0 * Bd * (lin - lin_last) * G),
        'tc_linear': ('tensor', 2.0 * Bd * (lin * P + (lin - lin_last) * G)),
        'sgemm_nn': ('tensor', 2.0 * Bd * lin * P), 'sgemm_nt': ('tensor', 2.0 * Bd * lin * G),
        'conv_tc_bwd': ('tensor', 2.0 * Bd * conv * P), 'conv_tc_fwd': ('tensor', 2.0 * Bd * conv * G),
        'conv_bwd': ('tensor', 2.0 * Bd * conv * P), 'conv_fwd': ('tensor', 2.0 * Bd * conv * G),
        # HBM-bound classes: bytes the class itself must move per sub-domain and launch set
        'relu_bwd': ('hbm', Bd * 20.0 * nr * P), 'relu_grad': ('hbm', Bd * 28.0 * nr * G),
        'adam': ('hbm', Bd * 36.0 * nr * G), 'chan': ('hbm', Bd * 8.0 * nr * (P + G)),
        'elementwise': ('hbm', Bd * 8.0 * nr * (P + G)),
    }, n_first


def measure(args, dist, rank, local, world, workload, Bd, steps, warmup, full):
    """All legs of one workload; returns the dict that becomes the JSON line (headline) or a `workloads` entry."""
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to
    dev = torch.device('cuda', local)
    wl = synth.WORKLOADS[workload]
    nodes = synth.build_nodes(workload, seed=0)
    plan = capi.Plan(nodes_to(nodes, dev))
    work = synth.algorithmic_work(nodes)
    NB = 3 if Bd * work['n_relu'] * 40 < 6e9 else 2
    pool = [synth.make_batch(nodes, Bd, wl['eps'], seed=1000 * rank + j, device=dev, bounds=wl.get('bounds', 'ibp'))
            for j in range(NB)]
    work_bufs = [clone_batch(b) for b in pool]
    gathered = torch.empty(world * Bd, 1, device=dev) if dist is not None else None

    def reset(j):
        for a_w, a_0 in zip(work_bufs[j]['alpha'], pool[j]['alpha']):
            a_w.copy_(a_0)
        for b_w, b_0 in zip(work_bufs[j]['beta'], pool[j]['beta']):
            b_w['val'].copy_(b_0['val'])

    def step_f2(i):
        j = i % NB
        reset(j)
        w = work_bufs[j]
        lb, lA, _ = plan.optimize(w['C'], w['x_L'], w['x_U'], w['lower'], w['upper'], w['alpha'], None, w['beta'], None,
                                  iteration=ITERATION, early_stop=False, early_stop_patience=10 ** 6, want_lA=True)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, lb)
        return lb

    def step_f1(i):
        w = work_bufs[i % NB]
        lb, _ = plan.crown_pass(w['C'], w['x_L'], w['x_U'], w['lower'], w['upper'], w['alpha'], None, None, want_lA=False)
        return lb

    for i in range(warmup):
        step_f2(i)
        # the synthetic split indices were range-checked by the first warm-up call; the device-resident legs below
        # time the kernels, not that check (one reduction + sync per call, kept on in the host-buffer legs)
        plan.validate_indices = False
    sampler = None
    if rank == 0 and full:
        sampler = NvmlSampler(local)
        if not sampler.start():
            sampler = ClockSampler(local)
            sampler.start()
    # the timed region: K steps, no per-launch instrumentation (the library's per-launch CUDA events cost ~5 us a launch,
    # 5 % of a step of ~1200 launches)
    l0 = capi.launch_count()
    sec = timed(dist, step_f2, steps)
    launches = capi.launch_count() - l0
    clocks = sampler.stop() if sampler is not None else None
    # the same steps again with a CUDA event pair around every launch: per-kernel-class times for the breakdown and the
    # roofline of the dominant class (`profiled_ms_per_step` says what the instrumentation costs)
    prof, psteps, sec_p = {}, max(2, min(steps, 4)), 0.0
    if not args.no_profile:
        capi.profile_enable(True)
        sec_p = timed(dist, step_f2, psteps)
        capi.profile_enable(False)
        prof = capi.profile_collect()
    value = world * Bd * steps / sec
    res = {'value': round(value, 1), 'ms_per_step': round(sec / steps * 1e3, 3), 'steps': steps, 'gpu_launches': int(launches),
           'config': config_of(workload, Bd, world), 'clocks': clocks,
           'plan': {'conv_on_tensor_cores': plan.conv_tc, 'conv_choices': plan.conv_choices, 'linear_on_tensor_cores': plan.tc_contractions, 'chain': plan.chain}}
    # ---- kernel breakdown and roofline of the dominant kernel class ---------------------------------------
    pk = peaks()
    if prof:
        classes, _ = kernel_classes(nodes, work, Bd)
        tot = sum(v['ms'] for v in prof.values())
        res['profiled_ms_per_step'] = round(sec_p / psteps * 1e3, 3)
        res['kernel_breakdown'] = {k: {'ms_per_step': round(v['ms'] / psteps, 4), 'launches_per_step': v['launches'] // psteps,
                                       'share': round(v['ms'] / tot, 4)}
                                   for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms'])}
        dom = max(prof.items(), key=lambda kv: kv[1]['ms'])[0]
        kind, amount = classes.get(dom, ('hbm', Bd * 16.0 * work['n_relu'] * ITERATION))
        per_step_ms, n_launch = prof[dom]['ms'] / psteps, prof[dom]['launches'] / psteps
        if kind == 'tensor':
            ach = amount / (per_step_ms * 1e-3) / 1e12
            peak = pk['bf16_tflops_sustained'] / 6.0
            res['roofline'] = {'kernel': dom, 'bound': 'tensor', 'achieved': round(ach, 3), 'peak': round(peak, 1),
                               'unit': 'TFLOP/s', 'frac': round(ach / peak, 4), 'frac_of_burst_peak': round(ach / (pk['bf16_tflops'] / 6.0), 4),
                               'traffic': None, 'launches_per_step': n_launch, 'avg_launch_us': round(per_step_ms * 1e3 / n_launch, 2),
                               'algorithmic_flop_per_launch': amount / n_launch,
                               'peak_source': f"{pk['source']} bf16_tflops_sustained / 6 (fp32-faithful bf16x3 split = 6 bf16 MMAs per product)"}
        else:
            ach = amount / (per_step_ms * 1e-3) / 1e9
            res['roofline'] = {'kernel': dom, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                               'frac': round(ach / pk['hbm_gbs'], 4), 'traffic': None, 'launches_per_step': n_launch,
                               'avg_launch_us': round(per_step_ms * 1e3 / n_launch, 2),
                               'algorithmic_bytes_per_launch': amount / n_launch, 'peak_source': pk['source']}
        try:
            with open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')) as f:
                tr = json.load(f)
            ent = tr.get(workload, {}).get(dom) if isinstance(tr.get(workload), dict) else None
            if ent and Bd == ent.get('bd'):
                res['roofline']['traffic'] = ent['bytes_per_launch']
                res['roofline']['traffic_source'] = ent['source']
        except Exception:
            pass
    bytes_f2 = (40.0 * work['n_relu'] + 8.0 * work['n_in']) * ITERATION
    flop_f2 = 2.0 * work['mac'] * (2 * ITERATION - 1)
    hbm_roof = pk['hbm_gbs'] * 1e9 / bytes_f2
    tc_roof = pk['bf16_tflops_sustained'] / 6.0 * 1e12 / flop_f2
    res['step_roofline'] = {'hbm_roof_subdomains_per_s': round(hbm_roof, 1), 'tensor_roof_subdomains_per_s': round(tc_roof, 1),
                            'frac_of_binding_roof': round(value / world / min(hbm_roof, tc_roof), 4),
                            'algorithmic_bytes_per_subdomain': bytes_f2, 'algorithmic_flop_per_subdomain': flop_f2}
    if args.only_f2:
        return res, nodes, pool

    # ---- F1, device-resident --------------------------------------------------------------------------
    for i in range(warmup):
        step_f1(i)
    n1 = steps * 4
    sec_f1 = timed(dist, step_f1, n1)
    res['f1'] = {'value': round(world * Bd * n1 / sec_f1, 1), 'unit': UNIT, 'ms_per_step': round(sec_f1 / n1 * 1e3, 4),
                 'what': 'one CROWN pass per sub-domain (reuse_alpha), device-resident'}

    # ---- e2e: the BaB step on the device-resident store; decisions from pinned host memory ----------------
    del work_bufs
    torch.cuda.empty_cache()
    res['e2e'] = e2e_device_store(args, dist, rank, world, workload, nodes, plan, pool[0], Bd, steps, warmup, gathered, res, full)
    if full:
        try:
            res['e2e_host_buffers'] = e2e_host_buffers(dist, world, plan, pool, Bd, steps, warmup, gathered)
        except RuntimeError as e:           # the pinned copies of a large workload may not fit the host
            res['e2e_host_buffers'] = {'unavailable': str(e)[:120]}
    return res, nodes, pool


def e2e_device_store(args, dist, rank, world, workload, nodes, plan, batch, Bd, steps, warmup, gathered, res, full):
    from types import SimpleNamespace
    from neuralsat_b200.abstractor import AbstractResults
    from neuralsat_b200.domain_store import DeviceBaB, DeviceDomainStore
    from neuralsat_b200.graph import activation_indices, preact_indices
    dev = plan.device
    B = Bd // 2
    acts, pres = activation_indices(nodes), preact_indices(nodes)

    def named(i):
        n = SimpleNamespace(name=nodes[i]['name'], index=i, op=nodes[i]['op'], output_shape=(1, *nodes[i]['shape']),
                            alpha_indices=None, inputs=[])
        return n
    act_nodes = [named(a) for a in acts]
    for a, p in zip(act_nodes, pres):
        a.inputs = [named(p)]
    graph_dev = plan.nodes
    net = SimpleNamespace(plan=plan, device=dev, final_name=nodes[-1]['name'], perturbed_optimizable_activations=act_nodes,
                          _dev_graph=lambda: graph_dev)
    sl = lambda t: t[:B]
    hist = None
    root = AbstractResults(
        objective_ids=torch.arange(B), output_lbs=torch.zeros(B, 1), rhs=torch.full((B, 1), float('inf')),
        cs=sl(batch['C']), input_lowers=sl(batch['x_L']), input_uppers=sl(batch['x_U']),
        lower_bounds={nodes[p]['name']: sl(batch['lower'][k]) for k, p in enumerate(pres)},
        upper_bounds={nodes[p]['name']: sl(batch['upper'][k]) for k, p in enumerate(pres)},
        lAs={nodes[a]['name']: torch.zeros(B, 1, *nodes[a]['shape'], device=dev) for a in acts},
        slopes={nodes[a]['name']: {nodes[-1]['name']: batch['alpha'][k][:, :, :B]} for k, a in enumerate(acts)},
        histories=hist)
    store = DeviceDomainStore(net, root, capacity=4 * B)
    # the parents carry the synthetic split histories of the batch (up to 16 records per layer)
    for k in range(len(acts)):
        bt = batch['beta'][k]
        J = bt['loc'].shape[1]
        store._grow_hist(k, J + 1)
        store.h_loc[k][:B, :J] = bt['loc'][:B].to(torch.int32)
        store.h_sign[k][:B, :J] = bt['sign'][:B]
        store.h_cnt[k][:B] = (bt['sign'][:B] != 0).sum(1).to(torch.int32)
        store.max_cnt[k] = int(store.h_cnt[k][:B].max())
    bab = DeviceBaB(net, store, iteration=ITERATION, early_stop=False, early_stop_patience=10 ** 6)
    g = torch.Generator().manual_seed(7 + rank)
    n_layers = len(acts)
    h_layer = torch.randint(0, n_layers, (B,), generator=g, dtype=torch.int32).pin_memory()
    sizes = torch.tensor(store.n_k)
    h_neuron = (torch.rand(B, generator=g) * sizes[h_layer.long()]).to(torch.int32).pin_memory()
    h_lb = torch.empty(Bd, 1).pin_memory()
    n0 = store.n
    max0 = list(store.max_cnt)

    dbg_t = []

    def step(i):
        if os.environ.get('CB_BENCH_DEBUG') == '1':
            dbg_t.append(time.perf_counter())
        store.n, store.max_cnt = n0, list(max0)           # the same parents every step: constant work
        dl = h_layer.to(dev, non_blocking=True)
        dn = h_neuron.to(dev, non_blocking=True)
        info = bab.step(B, decisions=(dl, dn))
        if dist is not None:
            dist.all_gather_into_tensor(gathered, bab.last['lb'])
        h_lb.copy_(bab.last['lb'], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return info

    for i in range(max(3, warmup)):
        info = step(i)
    assert info['kept'] == Bd, info
    if res.get('ms_per_step', 1e9) < 20.0:
        steps = max(steps, 20)                # millisecond steps: enough of them for a stable mean
    sec = timed(dist, step, steps)
    if dbg_t:
        print('e2e step starts (ms):', [round((b - a) * 1e3, 1) for a, b in zip(dbg_t, dbg_t[1:])], file=sys.stderr)
    out = {'value': round(world * Bd * steps / sec, 1), 'unit': UNIT, 'ms_per_step': round(sec / steps * 1e3, 3),
           'h2d_bytes_per_step': int(h_layer.numel() * 4 + h_neuron.numel() * 4),
           'd2h_bytes_per_step': int(h_lb.numel() * 4 + 4 * (1 + n_layers)),
           'what': 'DeviceBaB.step: split decisions from pinned host memory, children built / bounded / pruned / appended in '
                   'the device-resident domain store, lower bounds and survivor count read back'}
    if full:
        # branching leg: BaBSR scores + top-k + batched look-ahead passes + arg-max for B parents
        store.n = n0
        pick = store.pick_out(B)
        store.n = n0
        bab.topk = 10
        bab.branch(pick)
        torch.cuda.synchronize()
        reps = max(2, steps // 4)
        sec_b = timed(dist, lambda i: bab.branch(pick), reps)
        out_b = {'value': round(world * B * reps / sec_b, 1), 'unit': 'decisions/s', 'ms_per_call': round(sec_b / reps * 1e3, 3),
                 'what': f'BaBSR + top-10 look-ahead ({10 * 4 * B} CROWN passes per call) + arg-max for {B} parents, on the device'}
        res['branching'] = out_b
    return out


def e2e_host_buffers(dist, world, plan, pool, Bd, steps, warmup, gathered):
    from neuralsat_b200.pipeline import HostPipeline
    plan.validate_indices = True              # the call a user makes: range check of the split indices included

    def pin(t):
        return t.cpu().pin_memory()
    host = []
    for b in pool[:2]:
        host.append({'C': pin(b['C']), 'x_L': pin(b['x_L']), 'x_U': pin(b['x_U']),
                     'lower': [pin(t) for t in b['lower']], 'upper': [pin(t) for t in b['upper']],
                     # slopes are held in half precision on the host, as in the reference's domain store
                     'alpha': [pin(t.half()) for t in b['alpha']],
                     'beta': [{k: (None if v is None else pin(v)) for k, v in bt.items()} for bt in b['beta']]})
    gather = (lambda lb: dist.all_gather_into_tensor(gathered, lb)) if dist is not None else None
    DEPTH = 2
    pipe = HostPipeline(plan, depth=DEPTH, on_bounds=gather, iteration=ITERATION, early_stop=False, early_stop_patience=10 ** 6,
                        want_lA=True)

    def run_pipe(n):
        tickets = []
        for i in range(n):
            tickets.append(pipe.submit(host[i % 2]))
            if i >= DEPTH - 1:
                pipe.result(tickets[i - (DEPTH - 1)])
        for t in tickets[max(0, n - (DEPTH - 1)):]:
            pipe.result(t)
        pipe.drain()

    run_pipe(max(warmup, DEPTH + 1) + 2)
    sync_all(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    in0, out0 = pipe.total_in, pipe.total_out
    e0.record()
    run_pipe(steps)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    sec = float(ms.item()) / 1e3
    return {'value': round(world * Bd * steps / sec, 1), 'unit': UNIT, 'ms_per_step': round(sec / steps * 1e3, 3),
            'h2d_bytes_per_step': int((pipe.total_in - in0) // steps), 'd2h_bytes_per_step': int((pipe.total_out - out0) // steps),
            'what': 'HostPipeline.submit/result: every input from pinned host buffers, H2D / bounding / D2H of consecutive batches overlapped'}


def e2e_facade(workload, n_parents, steps, dev):
    """The reference-facing call itself: `neuralsat_b200.abstractor.NetworkAbstractor.forward(decisions, domain_params)`
    with HOST tensors in and out (per-domain history / beta dicts, fp16 slopes), exactly what the CPU arm times on the
    reference (oracle/ref_arm.py).  The per-domain Python lists of that API bound this leg, not the GPU."""
    from neuralsat_b200 import synth
    from neuralsat_b200.abstractor import AbstractResults, NetworkAbstractor
    wl = synth.WORKLOADS[workload]
    model = synth.build_network(workload, seed=0)
    ab = NetworkAbstractor(model, (1, *wl['in_shape']), 'crown-optimized', input_split=False, device=str(dev))
    net = ab.net
    B = n_parents
    b = synth.make_batch(net.graph, B, wl['eps'], seed=0, device='cpu', bounds=wl.get('bounds', 'ibp'))
    acts, pres, final = net.perturbed_optimizable_activations, net.split_nodes, net.final_name
    for m in acts:
        m.alpha = {final: torch.zeros(2, 1, 1, 1)}           # set_slope installs the stored slopes over this entry
    slopes = {m.name: {final: a.half()} for m, a in zip(acts, b['alpha'])}
    hist, betas = [], []
    for i in range(B):
        h, bt = {}, {}
        for p, rec in zip(pres, b['beta']):
            live = rec['sign'][i] != 0
            h[p.name] = (rec['loc'][i][live].clone(), rec['sign'][i][live].clone(), torch.zeros(int(live.sum())))
            bt[p.name] = rec['val'][i][live].clone()
        hist.append(h)
        betas.append(bt)
    g = torch.Generator().manual_seed(1)
    decisions = []
    for i in range(B):
        k = int(torch.randint(0, len(pres), (1,), generator=g))
        decisions.append([pres[k].name, int(torch.randint(0, b['lower'][k][0].numel(), (1,), generator=g)), 0.0])
    params = AbstractResults(objective_ids=torch.arange(B), output_lbs=torch.zeros(B, 1), input_lowers=b['x_L'], input_uppers=b['x_U'],
                             lower_bounds={p.name: l for p, l in zip(pres, b['lower'])},
                             upper_bounds={p.name: u for p, u in zip(pres, b['upper'])}, lAs=None, slopes=slopes, betas=betas,
                             histories=hist, cs=b['C'], rhs=torch.full((B, 1), float('inf')))
    net.set_bound_opts({'optimize_bound_args': {'early_stop_patience': 10 ** 6}})
    ab.forward(decisions, params)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = ab.forward(decisions, params)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {'value': round(2 * B / dt, 1), 'unit': UNIT, 'ms_per_call': round(dt * 1e3, 2), 'children_per_call': 2 * B,
            'what': 'NetworkAbstractor.forward(decisions, domain_params): host tensors and per-domain dicts in and out (the reference API)'}


def rebalance_leg(dist, rank, world, dev):
    """The work-queue exchange of the multi-GPU loop (shard.rebalance over NCCL) on unequal queues: rank r holds
    (r + 1) * 512 packed domain records of 64 KB; afterwards every rank holds the same number."""
    from neuralsat_b200 import shard
    counts = [(r + 1) * 512 for r in range(world)]
    n = counts[rank]
    rec = {'packed': torch.full((n, 16384), float(rank), device=dev), 'id': torch.arange(n, device=dev) + 100000 * rank}
    out = shard.rebalance(rec, counts, dist)                 # warm-up (communicator set-up)
    sync_all(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = shard.rebalance(rec, counts, dist)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    lens = torch.tensor([out['id'].shape[0]], device=dev)
    all_l = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(all_l, lens)
    after = [int(x) for x in all_l]
    moved = sum(max(0, c - a) for c, a in zip(counts, after))
    return {'ms': round(float(ms), 3), 'queues_before': counts, 'queues_after': after, 'records_moved': moved,
            'bytes_moved': moved * (16384 * 4 + 8), 'what': 'shard.rebalance: all_to_all_single of packed domain records over NCCL'}


def run_ours(args):
    dist, rank, local, world = dist_setup(args.gpus)
    Bd = args.bd or DEFAULT_BD[args.workload]
    res, nodes, pool = measure(args, dist, rank, local, world, args.workload, Bd, args.steps, args.warmup, full=True)
    if args.only_f2:
        if rank == 0:
            emit(json.dumps({'metric': METRIC, 'value': res['value'], 'unit': UNIT, 'ms_per_step': res['ms_per_step'],
                             'note': 'profiling run (--only-f2): not a bench line', 'plan': res.get('plan'), 'profiled_ms_per_step': res.get('profiled_ms_per_step'), 'kernel_breakdown': res.get('kernel_breakdown'),
                             'roofline': res.get('roofline')}))
        if dist is not None:
            dist.destroy_process_group()
        return
    cpu = None
    facade = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args, args.workload)
        try:
            facade = e2e_facade(args.workload, 512, 3, torch.device('cuda', local))
        except Exception as e:                                   # never lose the bench line to the compatibility leg
            facade = {'error': repr(e)[:200]}
    del pool
    torch.cuda.empty_cache()
    extras = {}
    if not args.no_extra:
        for w in EXTRA:
            if w == args.workload:
                continue
            try:
                r, _, p = measure(args, dist, rank, local, world, w, DEFAULT_BD[w], max(3, args.steps // 4), max(1, args.warmup // 2),
                                  full=False)
                del p
                extras[w] = {k: r[k] for k in ('value', 'ms_per_step', 'profiled_ms_per_step', 'steps', 'config', 'e2e', 'f1', 'roofline', 'step_roofline',
                                               'kernel_breakdown', 'plan') if k in r}
            except RuntimeError as e:
                extras[w] = {'error': str(e)[:200]}
            torch.cuda.empty_cache()
    reb = rebalance_leg(dist, rank, world, torch.device('cuda', local)) if dist is not None else None
    if rank == 0:
        line = {'metric': METRIC, 'value': res['value'], 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': res['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
                'data': 'synthetic', 'config': res['config'], 'clocks': res['clocks'], 'e2e': res['e2e'],
                'gpu_launches': res['gpu_launches'], 'e2e_host_buffers': res.get('e2e_host_buffers'), 'e2e_facade': facade, 'f1': res.get('f1'),
                'branching': res.get('branching'), 'roofline': res.get('roofline'), 'step_roofline': res['step_roofline'],
                'profiled_ms_per_step': res.get('profiled_ms_per_step'), 'kernel_breakdown': res.get('kernel_breakdown'), 'plan': res['plan'], 'cpu_baseline': cpu, 'workloads': extras,
                'rebalance': reb}
        emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def _oracle_inputs(nodes, b, n):
    from neuralsat_b200.graph import activation_indices, preact_indices
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    cpu = lambda t: t[:n].cpu() if t is not None else None
    lower = {p: cpu(b['lower'][k]) for k, p in enumerate(pres)}
    upper = {p: cpu(b['upper'][k]) for k, p in enumerate(pres)}
    alpha = {a: b['alpha'][k][:, :, :n].cpu() for k, a in enumerate(acts)}
    beta = {p: {kk: cpu(v) for kk, v in b['beta'][k].items()} for k, p in enumerate(pres)}
    return dict(C=cpu(b['C']), x_L=cpu(b['x_L']), x_U=cpu(b['x_U']), lower=lower, upper=upper, alpha=alpha,
                alpha_index={a: None for a in acts}, beta=beta)


class CpuArm:
    """The reference's CPU implementation of the path on all host cores: the UNMODIFIED reference when it has been
    staged (kind 'reference'), else its port (kind 'port').  One step = `n` children bounded (20 iterations)."""

    def __init__(self, workload, n):
        from oracle import ref_arm
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.n = n
        self.kind = 'port'
        if ref_arm.locate() is not None and os.environ.get('CB_BENCH_FORCE_PORT') != '1':
            try:
                self.arm = ref_arm.ReferenceArm(workload, max(1, n // 2))
                self.n = 2 * max(1, n // 2)
                self.kind = 'reference'
                self.step = self.arm.step
                self.sample = (f'each step = NetworkAbstractor.forward of the unmodified reference on {self.n // 2} parents = '
                               f'{self.n} children, {ITERATION} it., torch CPU fp32')
                return
            except Exception as e:               # staged tree unusable on this host: fall back to the port, and say so
                self.fallback_reason = repr(e)[:160]
        from neuralsat_b200 import synth
        from oracle import crown_oracle as orc
        torch.set_flush_denormal(True)
        wl = synth.WORKLOADS[workload]
        nodes = synth.build_nodes(workload, seed=0)
        batch = synth.make_batch(nodes, n, wl['eps'], seed=0, device='cpu', bounds=wl.get('bounds', 'ibp'))
        k = _oracle_inputs(nodes, batch, n)
        rhs = torch.full((n, 1), float('inf'))
        self.step = lambda: orc.optimize(nodes, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], k['alpha'], k['alpha_index'],
                                         k['beta'], rhs, iteration=ITERATION, early_stop_patience=10 ** 6)
        self.sample = f'each step = {n} sub-domains, {ITERATION} it. F2 step, oracle/crown_oracle.py (port), torch CPU fp32, flush-denormal on'

    def run(self, steps, warmup):
        for _ in range(warmup):
            self.step()
        t0 = time.perf_counter()
        for _ in range(steps):
            self.step()
        return (time.perf_counter() - t0) / steps


def cpu_baseline(args, workload):
    n = args.cpu_sample or CPU_SAMPLE[workload]
    arm = CpuArm(workload, n)
    dt = arm.run(3, 1)
    return {'value': round(arm.n / dt, 1), 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': arm.kind,
            'sample': arm.sample + f' ({dt:.2f} s per step, 3 steps after 1 warm-up)'}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the box's host cores, on the GPU arm's
    config (same JSON `config`), every step a bounded sample of it (cpu_baseline.sample)."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    world = max(1, args.gpus)
    Bd = args.bd or DEFAULT_BD[args.workload]
    n = args.cpu_sample or CPU_SAMPLE[args.workload]
    arm = CpuArm(args.workload, n)
    dt = arm.run(args.steps, max(1, min(args.warmup, 2)))
    v = round(arm.n / dt, 1)
    cpu = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': arm.kind, 'sample': arm.sample}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': round(dt * 1e3, 2), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic', 'config': config_of(args.workload, Bd, world),
            'cpu_baseline': cpu, 'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(json.dumps(line))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a CUDA device; the product path has no CPU fallback')
        run_ours(a)
