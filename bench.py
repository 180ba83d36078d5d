#!/usr/bin/env python
"""bench.py — BaB sub-domains bounded per second on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One *step* = one alpha/beta-CROWN bounding (form F2: 20 optimiser iterations, early stop disabled
so the work is constant) of one batch of Bd synthetic sub-domains per GPU of
the MNIST-FC 256x4 config (BASELINE.json configs[1]) through the C-ABI (libcrown_b200.so).
`value` = sub-domains bounded / s with inputs resident in HBM; `e2e` = same through the
plugin-level call with HOST (pinned) buffers, H2D and D2H inside the timed region.
Also reported: `f1` (one CROWN pass per domain, form F1), `roofline` of the dominant kernel
class (CUDA-event time per launch, measured inside the timed region by the library's profiler),
`cpu_baseline` (the CPU oracle = port of the reference, on the host cores, bounded sample).
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
    ap.add_argument('--workload', default='mnistfc_256x4')
    ap.add_argument('--bd', type=int, default=9472,
                    help='sub-domains per GPU per step (2*B children); default 148 SMs x 64-row tiles = one full wave')
    ap.add_argument('--cpu-sample', type=int, default=2048, help='sub-domains in the CPU baseline sample')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-profile', action='store_true')
    ap.add_argument('--only-f2', action='store_true', help='skip the F1 / e2e legs (profiling runs)')
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
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


def run_ours(args):
    from neuralsat_b200 import capi, synth
    from neuralsat_b200.graph import nodes_to, trace_module
    dist, rank, local, world = dist_setup(args.gpus)
    dev = torch.device('cuda', local)
    wl = synth.WORKLOADS[args.workload]
    nodes = synth.build_nodes(args.workload, seed=0)
    plan = capi.Plan(nodes_to(nodes, dev))
    work = synth.algorithmic_work(nodes)
    Bd = args.bd
    # pool of distinct input batches (each > L2 together with the workspace); rank-dependent seeds
    NB = 3
    pool = [synth.make_batch(nodes, Bd, wl['eps'], seed=1000 * rank + j, device=dev, bounds=wl.get('bounds', 'ibp')) for j in range(NB)]
    work_bufs = [clone_batch(b) for b in pool]
    gathered = torch.empty(world * Bd, 1, device=dev) if dist is not None else None

    def reset(j):
        # a BaB iteration receives fresh alpha/beta from the domain store: restore the parameters
        for a_w, a_0 in zip(work_bufs[j]['alpha'], pool[j]['alpha']):
            a_w.copy_(a_0)
        for b_w, b_0 in zip(work_bufs[j]['beta'], pool[j]['beta']):
            b_w['val'].copy_(b_0['val'])

    def step_f2(i):
        j = i % NB
        reset(j)
        w = work_bufs[j]
        lb, lA, _ = plan.optimize(w['C'], w['x_L'], w['x_U'], w['lower'], w['upper'], w['alpha'], None,
                                  w['beta'], None, iteration=ITERATION, early_stop=False,
                                  early_stop_patience=10 ** 6, want_lA=True)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, lb)     # per-domain lower bounds to every rank
        return lb

    def step_f1(i):
        w = work_bufs[i % NB]
        lb, _ = plan.crown_pass(w['C'], w['x_L'], w['x_U'], w['lower'], w['upper'], w['alpha'], None,
                                None, want_lA=False)
        return lb

    # ---- F2, device-resident -------------------------------------------------------------
    for i in range(args.warmup):
        step_f2(i)
    sampler = NvmlSampler(local)
    if rank == 0 and not sampler.start():
        sampler = ClockSampler(local)
        sampler.start()
    if not args.no_profile:
        capi.profile_enable(True)
    l0 = capi.launch_count()
    sec = timed(dist, step_f2, args.steps)
    launches = capi.launch_count() - l0
    capi.profile_enable(False)
    prof = capi.profile_collect()
    clocks = sampler.stop() if rank == 0 else None
    value = world * Bd * args.steps / sec
    if args.only_f2:
        if rank == 0:
            emit(json.dumps({'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'ms_per_step': round(sec / args.steps * 1e3, 3),
                              'note': 'profiling run (--only-f2): not a bench line', 'kernel_breakdown': prof}))
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- F1, device-resident -------------------------------------------------------------
    for i in range(args.warmup):
        step_f1(i)
    sec_f1 = timed(dist, step_f1, args.steps * 4)
    value_f1 = world * Bd * args.steps * 4 / sec_f1

    # ---- e2e: host (pinned) buffers, H2D + call + D2H inside the timed region --------------
    def pin(t):
        return t.cpu().pin_memory()
    host = []
    for b in pool[:2]:
        host.append({'C': pin(b['C']), 'x_L': pin(b['x_L']), 'x_U': pin(b['x_U']),
                     'lower': [pin(t) for t in b['lower']], 'upper': [pin(t) for t in b['upper']],
                     # slopes are held in half precision on the host, as in the reference's domain store
                     # (get_slope(half=True), NS/abstractor/utils.py:51-59; the synthetic values are fp16-exact)
                     'alpha': [pin(t.half()) for t in b['alpha']],
                     'beta': [{k: (None if v is None else pin(v)) for k, v in bt.items()} for bt in b['beta']]})
    h2d = sum(t.numel() * t.element_size() for h in host[:1] for t in
              [h['C'], h['x_L'], h['x_U']] + h['lower'] + h['upper'] + h['alpha'] +
              [v for bt in h['beta'] for v in bt.values() if v is not None])
    out_host = {}

    def step_e2e(i):
        h = host[i % 2]
        nb = True
        d = {'C': h['C'].to(dev, non_blocking=nb), 'x_L': h['x_L'].to(dev, non_blocking=nb),
             'x_U': h['x_U'].to(dev, non_blocking=nb),
             'lower': [t.to(dev, non_blocking=nb) for t in h['lower']],
             'upper': [t.to(dev, non_blocking=nb) for t in h['upper']],
             'alpha': [t.to(dev, non_blocking=nb).float() for t in h['alpha']],
             'beta': [{k: (None if v is None else v.to(dev, non_blocking=nb)) for k, v in bt.items()}
                      for bt in h['beta']]}
        lb, lA, _ = plan.optimize(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'], None,
                                  d['beta'], None, iteration=ITERATION, early_stop=False,
                                  early_stop_patience=10 ** 6, want_lA=True)
        if dist is not None:
            dist.all_gather_into_tensor(gathered, lb)
        # what NetworkAbstractor._forward_hidden returns to the host (abstractor.py:315-323):
        # lbs, lAs, slopes as fp16, betas
        out_host['lb'] = lb.to('cpu', non_blocking=True)
        out_host['lA'] = [t.to('cpu', non_blocking=True) for t in lA]
        out_host['alpha'] = [t.half().to('cpu', non_blocking=True) for t in d['alpha']]
        out_host['beta'] = [bt['val'].to('cpu', non_blocking=True) for bt in d['beta']]
        torch.cuda.current_stream().synchronize()

    for i in range(2):
        step_e2e(i)
    sec_e2e_sync = timed(dist, step_e2e, args.steps)
    d2h = sum(t.numel() * t.element_size() for t in
              [out_host['lb']] + out_host['lA'] + out_host['alpha'] + out_host['beta'])

    # the same through the public host-buffer pipeline (neuralsat_b200.pipeline.HostPipeline): H2D of batch
    # i+1 and D2H of batch i-1 overlap the bounding of batch i; every step still copies all of its inputs
    # from pinned host memory and all of its results back, inside the timed region
    from neuralsat_b200.pipeline import HostPipeline
    gather = (lambda lb: dist.all_gather_into_tensor(gathered, lb)) if dist is not None else None
    DEPTH = int(os.environ.get('CB_PIPE_DEPTH', '2'))       # batches in flight (results are read DEPTH - 1 submissions later); 3 and 4 measured slower
    pipe = HostPipeline(plan, depth=DEPTH, on_bounds=gather, iteration=ITERATION, early_stop=False, early_stop_patience=10 ** 6, want_lA=True)

    def run_pipe(n):
        tickets = []
        for i in range(n):
            tickets.append(pipe.submit(host[i % 2]))
            if i >= DEPTH - 1:
                pipe.result(tickets[i - (DEPTH - 1)])
        for t in tickets[max(0, n - (DEPTH - 1)):]:
            pipe.result(t)
        pipe.drain()

    run_pipe(max(args.warmup, DEPTH + 1) + 3)          # the allocator needs a few batches to settle its cross-stream reuse
    sync_all(dist)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    in0, out0 = pipe.total_in, pipe.total_out
    e0.record()
    run_pipe(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device='cuda')
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.barrier()
    sec_e2e = float(ms.item()) / 1e3
    h2d = (pipe.total_in - in0) // args.steps          # counted from the tensors copied inside the timed region
    d2h = (pipe.total_out - out0) // args.steps
    value_e2e = world * Bd * args.steps / sec_e2e
    value_e2e_sync = world * Bd * args.steps / sec_e2e_sync

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel class ---------------------------------------------
    pk = peaks()
    roofline = None
    breakdown = {}
    if prof:
        tot_ms = sum(v['ms'] for v in prof.values())
        for k, v in sorted(prof.items(), key=lambda kv: -kv[1]['ms']):
            breakdown[k] = {'ms_per_step': round(v['ms'] / args.steps, 4), 'launches_per_step': v['launches'] // args.steps,
                            'share': round(v['ms'] / tot_ms, 4)}
        dom = max(prof.items(), key=lambda kv: kv[1]['ms'])[0]
        dims = [(nd['weight'].shape[0], nd['weight'].shape[1]) for nd in nodes if nd['op'] == 'linear']
        flops_pass = 2.0 * Bd * sum(o * i for o, i in dims)
        flops_grad = 2.0 * Bd * sum(o * i for o, i in dims[:-1])
        # fp32-faithful tensor peak: TF32 runs at half the bf16 rate and a 3xTF32 split needs
        # 3 MMAs per product -> measured bf16 / 6 (TF32 peak itself is not in MEASURED_PEAKS.json)
        tc_peak = pk['bf16_tflops_sustained'] / 6.0
        per_step_ms = prof[dom]['ms'] / args.steps
        n_launch = prof[dom]['launches'] / args.steps
        chain = bool(getattr(plan, 'chain', False))
        tensor_flops = {'sgemm_nn': flops_pass * ITERATION, 'sgemm_nt': flops_grad * (ITERATION - 1),
                        # with the whole-network pass kernel the per-layer launches only do the gradient direction
                        'tc_linear': flops_grad * (ITERATION - 1) + (0.0 if chain else flops_pass * ITERATION),
                        'chain_pass': flops_pass * ITERATION, 'chain_grad': flops_grad * (ITERATION - 1)}
        if dom in tensor_flops:
            fl = tensor_flops[dom]
            ach = fl / (per_step_ms * 1e-3) / 1e12
            roofline = {'kernel': dom, 'bound': 'tensor', 'achieved': round(ach, 3), 'peak': round(tc_peak, 1),
                        'unit': 'TFLOP/s', 'frac': round(ach / tc_peak, 4), 'traffic': None,
                        'peak_source': f"{pk['source']} bf16_tflops_sustained / 6 (fp32-faithful bf16x3 split = 6 bf16 MMAs per product)",
                        'launches_per_step': n_launch, 'avg_launch_us': round(per_step_ms * 1e3 / n_launch, 2),
                        'algorithmic_flop_per_launch': fl / n_launch}
        else:
            # HBM-bound classes: algorithmic bytes of the whole F2 step (40*N_relu + 8*N_in per
            # domain and iteration) attributed to the class by its share would be meaningless; report bytes
            # the class itself must move
            by = Bd * (16.0 * work['n_relu']) * ITERATION
            ach = by / (per_step_ms * 1e-3) / 1e9
            roofline = {'kernel': dom, 'bound': 'hbm', 'achieved': round(ach, 1), 'peak': pk['hbm_gbs'],
                        'unit': 'GB/s', 'frac': round(ach / pk['hbm_gbs'], 4), 'traffic': None,
                        'peak_source': pk['source'], 'launches_per_step': n_launch}
    # DRAM traffic of the dominant kernel from the committed ncu capture (per launch), if one exists
    try:
        tr = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if roofline is not None and roofline['kernel'] in tr and args.workload == 'mnistfc_256x4' and Bd == tr.get('_bd', 8192):
            roofline['traffic'] = tr[roofline['kernel']]['bytes_per_launch']
            roofline['traffic_source'] = tr[roofline['kernel']]['source']
    except Exception:
        pass
    # whole-step HBM roofline: bytes_F2_iter ~ 40*N_relu + 8*N_in per domain
    bytes_f2 = (40.0 * work['n_relu'] + 8.0 * work['n_in']) * ITERATION
    step_roof = {'hbm_roof_subdomains_per_s': round(pk['hbm_gbs'] * 1e9 / bytes_f2, 1),
                 'frac_of_hbm_roof': round(value / world / (pk['hbm_gbs'] * 1e9 / bytes_f2), 4),
                 'algorithmic_bytes_per_subdomain': bytes_f2}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baseline(args, nodes, pool[0])

    line = {
        'metric': METRIC, 'value': round(value, 1), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': round(sec / args.steps * 1e3, 3), 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.workload} eps={wl["eps"]} F2 alpha/beta-CROWN step ({ITERATION} it., early stop off)',
                   'subdomains_per_gpu_per_step': Bd, 'spec_rows': 1, 'l2_policy': f'{NB} rotating input batches, working set > L2',
                   'parallelism': f'domains sharded x{world}, lb all_gather per step' if world > 1 else 'single GPU'},
        'clocks': clocks,
        'e2e': {'value': round(value_e2e, 1), 'unit': UNIT, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                'ms_per_step': round(sec_e2e / args.steps * 1e3, 3),
                'what': 'HostPipeline.submit/result: pinned host buffers, H2D / bounding / D2H of consecutive batches overlapped',
                'unpipelined': {'value': round(value_e2e_sync, 1), 'ms_per_step': round(sec_e2e_sync / args.steps * 1e3, 3),
                                'what': 'one blocking call per batch: H2D, Plan.optimize, D2H, stream sync'}},
        'gpu_launches': int(launches),
        'f1': {'value': round(value_f1, 1), 'unit': UNIT, 'ms_per_step': round(sec_f1 / (args.steps * 4) * 1e3, 4),
               'what': 'one CROWN pass per sub-domain (reuse_alpha), device-resident'},
        'roofline': roofline, 'step_roofline': step_roof, 'kernel_breakdown': breakdown,
        'cpu_baseline': cpu,
    }
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


def cpu_time_sample(nodes_cpu, k, n, reps=1):
    """The CPU oracle (port of auto_LiRPA's path) timed on n sub-domains, all host threads."""
    from oracle import crown_oracle as orc
    rhs = torch.full((n, 1), float('inf'))
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.optimize(nodes_cpu, k['C'], k['x_L'], k['x_U'], k['lower'], k['upper'], k['alpha'],
                     k['alpha_index'], k['beta'], rhs, iteration=ITERATION, early_stop_patience=10 ** 6)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def cpu_baseline(args, nodes, batch):
    from neuralsat_b200.graph import nodes_to
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_flush_denormal(True)      # favours the CPU: the reference itself runs with denormals on
    n = min(args.cpu_sample, args.bd)
    nodes_cpu = nodes_to(nodes, 'cpu')
    k = _oracle_inputs(nodes, batch, n)
    cpu_time_sample(nodes_cpu, _oracle_inputs(nodes, batch, min(256, n)), min(256, n))   # warm-up
    dt = cpu_time_sample(nodes_cpu, k, n, reps=2)
    return {'value': round(n / dt, 1), 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': f'{n} sub-domains of the same workload, {ITERATION} it. F2 step, best of 2, '
                      f'torch CPU fp32 with flush-denormal on ({dt:.2f} s)'}


def run_reference(args):
    """Reference arm: the reference's CPU implementation of the path.  /root/reference is Python and
    cannot travel to the GPU box, so this times the CPU oracle (its port, pinned bit-exact to the
    reference by tests/test_oracle_golden.py) on the box's host cores."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from neuralsat_b200 import synth
    from neuralsat_b200.graph import trace_module
    wl = synth.WORKLOADS[args.workload]
    nodes = synth.build_nodes(args.workload, seed=0)
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.set_flush_denormal(True)
    n = args.cpu_sample
    batch = synth.make_batch(nodes, n, wl['eps'], seed=0, device='cpu')
    k = _oracle_inputs(nodes, batch, n)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_time_sample(nodes, _oracle_inputs(nodes, batch, min(256, n)), min(256, n))
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_time_sample(nodes, k, n)
    dt = (time.perf_counter() - t0) / steps
    v = round(n / dt, 1)
    cpu = {'value': v, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
           'sample': f'each step = {n} sub-domains, {ITERATION} it. F2 step, torch CPU fp32, flush-denormal on'}
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
            'warmup': args.warmup, 'ms_per_step': round(dt * 1e3, 2), 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': f'{args.workload} eps={wl["eps"]} F2 alpha/beta-CROWN step ({ITERATION} it., early stop off)',
                       'subdomains_per_step': n},
            'cpu_baseline': cpu,
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(json.dumps(line))


if __name__ == '__main__':
    a = parse()
    if a.impl == 'reference':
        run_reference(a)
    else:
        if not torch.cuda.is_available():
            raise RuntimeError('bench.py needs a CUDA device; the product path has no CPU fallback')
        run_ours(a)
