"""Host-buffer pipeline around `capi.Plan.optimize` (form F2, NS/abstractor/abstractor.py:289-323).

The reference moves every picked batch host -> device before bounding it and the results device -> host
afterwards (`DomainsList.pick_out(batch, device)`, NS/heuristic/domains_list.py:153-227;
`AbstractResults` tensors go back with `.to('cpu')`, NS/abstractor/abstractor.py:315-344), strictly one after
the other.  Here the three phases of consecutive batches overlap: the inputs of batch i+1 cross PCIe on a copy
stream while batch i is being bounded, and the results of batch i-1 return on a third stream.  Device buffers
are double-buffered per slot; all ordering is by CUDA events, the host only blocks in `result()`.

    pipe = HostPipeline(plan)
    t0 = pipe.submit(host_batch0)          # pinned host tensors: C, x_L, x_U, lower[], upper[], alpha[], beta[]
                                           # alpha[] may be fp16 - the reference's host domain store holds the slopes in
                                           # half precision (get_slope(half=True), NS/abstractor/utils.py:51-59) - and
                                           # is then widened to fp32 on the device after the copy
    t1 = pipe.submit(host_batch1)
    out0 = pipe.result(t0)                 # {'lb', 'lA'[], 'alpha'[] (fp16, as get_slope), 'beta'[]} on the host
"""
from __future__ import annotations

from typing import Dict, List, Optional

import torch


def _pin_like(t: torch.Tensor, dtype=None) -> torch.Tensor:
    return torch.empty(t.shape, dtype=dtype or t.dtype, device='cpu').pin_memory()


class _Slot:
    def __init__(self):
        self.dev: Optional[dict] = None        # device input buffers
        self.host_out: Optional[dict] = None   # pinned host output buffers
        self.ev_compute = None                 # bounding of the batch in this slot has finished
        self.ev_out = None                     # its results are on the host
        self.keep = None                       # device outputs kept alive until the D2H has finished
        self.bytes_in = 0
        self.bytes_out = 0
        self.sig = None                        # shapes and dtypes of the host batch the buffers were made for


class HostPipeline:
    def __init__(self, plan, depth: int = 2, on_bounds=None, **optimize_kwargs):
        """on_bounds(lb_device): called on the compute stream right after a batch has been enqueued, e.g. the
        all_gather of the per-domain lower bounds across ranks (shard.gather_lower_bounds)."""
        self.plan = plan
        self.on_bounds = on_bounds
        self.device = plan.device
        self.kw = optimize_kwargs
        self.slots = [_Slot() for _ in range(depth)]
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self.n = 0
        self.total_in = 0          # bytes copied host -> device / device -> host so far
        self.total_out = 0

    @staticmethod
    def _flat(b: dict) -> List[torch.Tensor]:
        out = [b['C'], b['x_L'], b['x_U']] + list(b['lower']) + list(b['upper']) + list(b['alpha'])
        if b.get('rhs') is not None:
            out.append(b['rhs'])
        for bt in b.get('beta') or []:
            out += [v for v in bt.values() if v is not None]
        return out

    def _alloc_like(self, host: dict) -> dict:
        dev = self.device
        mk = lambda t: torch.empty(t.shape, dtype=t.dtype, device=dev)
        mk32 = lambda t: torch.empty(t.shape, dtype=torch.float32, device=dev)
        d = {'C': mk(host['C']), 'x_L': mk(host['x_L']), 'x_U': mk(host['x_U']),
             'lower': [mk(t) for t in host['lower']], 'upper': [mk(t) for t in host['upper']],
             'alpha': [mk32(t) for t in host['alpha']], 'beta': None,
             'rhs': None if host.get('rhs') is None else mk(host['rhs']),
             # fp16 slopes land in a staging buffer and are widened into 'alpha' on the copy stream
             'alpha16': [mk(t) if t.dtype == torch.float16 else None for t in host['alpha']]}
        if host.get('beta') is not None:
            d['beta'] = [{k: (None if v is None else mk(v)) for k, v in bt.items()} for bt in host['beta']]
        return d

    def submit(self, host: dict) -> int:
        """Enqueue H2D -> bounding -> D2H of one batch of pinned host tensors; returns a ticket."""
        ticket = self.n
        self.n += 1
        slot = self.slots[ticket % len(self.slots)]
        main = torch.cuda.current_stream(self.device)
        sig = [(t.shape, t.dtype) for t in self._flat(host)]
        if slot.dev is None or slot.sig != sig:
            # shapes changed (batch size, beta records per layer): the old buffers may still be in use on the
            # copy streams, so let the slot run dry before they go back to the allocator
            for ev in (slot.ev_compute, slot.ev_out):
                if ev is not None:
                    ev.synchronize()
            slot.dev = self._alloc_like(host)
            slot.sig = sig
            slot.host_out = None
        # the slot's device buffers are free once its previous batch has been bounded and read back
        if slot.ev_compute is not None:
            self.h2d.wait_event(slot.ev_compute)
        if slot.ev_out is not None:
            self.h2d.wait_event(slot.ev_out)
        with torch.cuda.stream(self.h2d):
            for dst, src in zip(self._flat(slot.dev), self._flat(host)):
                if dst.dtype == src.dtype:
                    dst.copy_(src, non_blocking=True)
            for dst, stage, src in zip(slot.dev['alpha'], slot.dev['alpha16'], host['alpha']):
                if stage is not None:
                    stage.copy_(src, non_blocking=True)
                    dst.copy_(stage)                              # fp16 -> fp32 on the device
            ev_in = self.h2d.record_event()
        slot.bytes_in = sum(t.numel() * t.element_size() for t in self._flat(host))
        main.wait_event(ev_in)
        d = slot.dev
        # rhs: the decision threshold of the stop criterion (without it nothing is ever "verified" and every domain
        # keeps optimising to the last iteration); alpha_pos: device int32 maps of sparse-feature slopes.  The
        # H2D / bounding overlap needs early_stop=False (the exact early exit synchronises the stream per iteration).
        # the split indices are range-checked on the host copy (no stream sync, which would undo the overlap)
        if self.plan.validate_indices and host.get('beta') is not None:
            for k, bt in enumerate(host['beta']):
                if bt is not None and bt['loc'].numel() and (int(bt['loc'].min()) < 0 or
                                                              int(bt['loc'].max()) >= self.plan.act_numel[k]):
                    raise ValueError('beta loc holds a neuron index outside its layer')
        checked, self.plan.validate_indices = self.plan.validate_indices, False
        try:
            lb, lA, _ = self.plan.optimize(d['C'], d['x_L'], d['x_U'], d['lower'], d['upper'], d['alpha'],
                                           host.get('alpha_pos'), d['beta'], d.get('rhs'), **self.kw)
        finally:
            self.plan.validate_indices = checked
        if self.on_bounds is not None:
            self.on_bounds(lb)
        slot.ev_compute = main.record_event()
        if slot.host_out is None:
            slot.host_out = {'lb': _pin_like(lb), 'lA': [_pin_like(t) for t in (lA or [])],
                             'alpha': [_pin_like(t, torch.float16) for t in d['alpha']],
                             'beta': [_pin_like(bt['val']) for bt in (d['beta'] or [])]}
        ho = slot.host_out
        with torch.cuda.stream(self.d2h):
            self.d2h.wait_event(slot.ev_compute)
            ho['lb'].copy_(lb, non_blocking=True)
            for dst, src in zip(ho['lA'], lA or []):
                dst.copy_(src, non_blocking=True)
            # slopes travel as fp16 (NS/abstractor/utils.py:51-59); the device-side fp16 buffers belong to the slot (no
            # per-batch allocation on a side stream: the caching allocator would keep growing until it has seen every
            # cross-stream reuse pattern)
            if d.get('alpha_out16') is None:
                d['alpha_out16'] = [torch.empty(t.shape, dtype=torch.float16, device=self.device) for t in d['alpha']]
            halves = d['alpha_out16']
            for buf, src in zip(halves, d['alpha']):
                buf.copy_(src)
            for dst, src in zip(ho['alpha'], halves):
                dst.copy_(src, non_blocking=True)
            for dst, bt in zip(ho['beta'], d['beta'] or []):
                dst.copy_(bt['val'], non_blocking=True)
            slot.ev_out = self.d2h.record_event()
        slot.keep = (lb, lA)
        for t in [lb] + list(lA or []):
            t.record_stream(self.d2h)
        slot.bytes_out = sum(t.numel() * t.element_size() for t in [ho['lb']] + ho['lA'] + ho['alpha'] + ho['beta'])
        self.total_in += slot.bytes_in
        self.total_out += slot.bytes_out
        return ticket

    def result(self, ticket: int) -> Dict[str, object]:
        """Blocks until the results of `ticket` are in pinned host memory and returns them (the buffers are
        reused `depth` submissions later)."""
        slot = self.slots[ticket % len(self.slots)]
        slot.ev_out.synchronize()
        return slot.host_out

    def drain(self) -> None:
        """Make the current stream wait for every outstanding read-back (so that an event recorded next on
        it brackets the whole pipeline)."""
        main = torch.cuda.current_stream(self.device)
        for s in self.slots:
            if s.ev_out is not None:
                main.wait_event(s.ev_out)
