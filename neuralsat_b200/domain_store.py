"""Device-resident domain store, branching and BaB step.

Mirrors, with every per-domain record kept in HBM:
  DomainsList.pick_out / add / __len__ / minimum_lowers ... NS/heuristic/domains_list.py:153-311
  TensorStorage (LIFO pop / append, geometric growth) ...... NS/util/misc/tensor_storage.py:4-97
  NetworkAbstractor._forward_hidden child construction ..... NS/abstractor/abstractor.py:253-298,
      hidden_split_idx, update_histories, set_beta ......... NS/abstractor/utils.py:159-250
  get_slope(half=True) fp16 slopes .......................... NS/abstractor/utils.py:51-59
  DecisionHeuristic.smart_hidden_branching / get_topk_scores  NS/heuristic/decision_heuristics.py:78-251
  _compute_babsr_scores ...................................... NS/heuristic/util.py:31-72

The reference moves every picked domain H2D and every child D2H (16 N_relu + ... bytes per domain and iteration) and
builds children, histories and betas in per-domain Python loops.  Here a BaB iteration is:
    pick (views of the last B records) -> BaBSR scores, top-k, ONE batched look-ahead pass, decision   [kernels]
    -> children by row gather + split kernel -> alpha/beta-CROWN (cb_optimize) -> keep / rank / append  [kernels]
and the host sees one small read-back per iteration (number of surviving children, history lengths).
ReLU networks with one spec-shared slope set per domain (S1 = 1), the regime of every BaB config in BASELINE.json.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional

import torch

from . import capi
from .capi import CbCopyDesc, CbSplitLayer


def _descs_to_device(descs: List[CbCopyDesc], device) -> torch.Tensor:
    arr = (CbCopyDesc * len(descs))(*descs)
    buf = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
    return buf.to(device)


def _desc(src: torch.Tensor, dst: torch.Tensor, width: int, mode: int, src_stride=None, dst_stride=None, dst_width=0,
          src_S=0, src_Bd=0) -> CbCopyDesc:
    d = CbCopyDesc()
    d.src, d.dst = src.data_ptr(), dst.data_ptr()
    d.width, d.dst_width, d.mode = int(width), int(dst_width), int(mode)
    d.src_S, d.src_Bd = int(src_S), int(src_Bd)
    d.src_stride = int(width if src_stride is None else src_stride)
    d.dst_stride = int(width if dst_stride is None else dst_stride)
    return d


class DeviceDomainStore:
    """Unverified sub-domains of one verification run as packed device records (LIFO, like the reference's store)."""

    def __init__(self, net, root, capacity: int = 4096):
        """net: neuralsat_b200.BoundedModule; root: AbstractResults of `initialize` (tensors on any device)."""
        self.net = net
        self.dev = net.device
        self.final_name = net.final_name
        self.acts = list(net.perturbed_optimizable_activations)
        self.pres = [a.inputs[0] for a in self.acts]
        if any(a.op != 'relu' for a in self.acts):
            raise NotImplementedError('the device store handles ReLU split points only')
        self.n_layers = len(self.acts)
        self.visited = 0
        dev = self.dev
        keep = (root.output_lbs.detach().cpu() <= root.rhs.detach().cpu()).all(1).nonzero().flatten()
        n = int(keep.numel())
        self.S = int(root.cs.shape[1])
        self.n_out = int(root.cs.shape[2])
        self.in_shape = tuple(root.input_lowers.shape[1:])
        self.n_in = int(root.input_lowers[0].numel())
        self.n_k = [int(root.lower_bounds[p.name][0].numel()) for p in self.pres]
        self.shape_k = [tuple(root.lower_bounds[p.name].shape[1:]) for p in self.pres]
        self.n_alpha = []
        for a in self.acts:
            sl = root.slopes[a.name][self.final_name]
            if sl.shape[1] != 1:
                raise NotImplementedError('per-spec slopes (S1 > 1) are not held by the device store')
            self.n_alpha.append(int(sl[0, 0, 0].numel()))
        self.cap = max(capacity, 2 * n)
        self.Jc = [4] * self.n_layers
        f32, kw = torch.float32, dict(device=dev)
        self.ids = torch.zeros(self.cap, dtype=torch.int64, **kw)
        self.lb = torch.zeros(self.cap, self.S, dtype=f32, **kw)
        self.cs = torch.zeros(self.cap, self.S * self.n_out, dtype=f32, **kw)
        self.rhs = torch.zeros(self.cap, self.S, dtype=f32, **kw)
        self.x_L = torch.zeros(self.cap, self.n_in, dtype=f32, **kw)
        self.x_U = torch.zeros(self.cap, self.n_in, dtype=f32, **kw)
        self.lower = [torch.zeros(self.cap, nk, dtype=f32, **kw) for nk in self.n_k]
        self.upper = [torch.zeros(self.cap, nk, dtype=f32, **kw) for nk in self.n_k]
        self.alpha = [torch.zeros(self.cap, na, dtype=torch.float16, **kw) for na in self.n_alpha]
        self.lA = [torch.zeros(self.cap, self.S * nk, dtype=f32, **kw) for nk in self.n_k]
        self.h_cnt = [torch.zeros(self.cap, dtype=torch.int32, **kw) for _ in range(self.n_layers)]
        self.h_loc = [torch.zeros(self.cap, J, dtype=torch.int32, **kw) for J in self.Jc]
        self.h_sign = [torch.zeros(self.cap, J, dtype=f32, **kw) for J in self.Jc]
        self.h_beta = [torch.zeros(self.cap, J, dtype=f32, **kw) for J in self.Jc]
        self.max_cnt = [0] * self.n_layers
        self.n = n
        if n:
            k = keep.to(root.output_lbs.device) if root.output_lbs.device.type != 'cpu' else keep
            sel = lambda t: t.detach()[keep.to(t.device)].to(dev)
            self.ids[:n] = sel(root.objective_ids)
            self.lb[:n] = sel(root.output_lbs)
            self.cs[:n] = sel(root.cs).reshape(n, -1)
            self.rhs[:n] = sel(root.rhs)
            self.x_L[:n] = sel(root.input_lowers).reshape(n, -1)
            self.x_U[:n] = sel(root.input_uppers).reshape(n, -1)
            for i, (a, p) in enumerate(zip(self.acts, self.pres)):
                self.lower[i][:n] = sel(root.lower_bounds[p.name]).reshape(n, -1)
                self.upper[i][:n] = sel(root.upper_bounds[p.name]).reshape(n, -1)
                sl = root.slopes[a.name][self.final_name].detach()
                self.alpha[i][:n] = sl[0, 0][keep.to(sl.device)].reshape(n, -1).to(dev).half()
                self.lA[i][:n] = sel(root.lAs[a.name]).reshape(n, -1)
            # split histories / betas of domains that are not roots (resuming from a host-side DomainsList)
            hist = root.histories if isinstance(root.histories, (list, tuple)) else None
            if hist is not None:
                for row, j in enumerate(keep.tolist()):
                    for i, p in enumerate(self.pres):
                        loc, sign = hist[j][p.name][:2]
                        c = len(loc)
                        if c == 0:
                            continue
                        self._grow_hist(i, c + 1)
                        self.h_cnt[i][row] = c
                        self.h_loc[i][row, :c] = torch.as_tensor(loc, dtype=torch.int32).to(dev)
                        self.h_sign[i][row, :c] = torch.as_tensor(sign, dtype=f32).to(dev)
                        if root.betas is not None and root.betas[j] is not None and p.name in root.betas[j]:
                            self.h_beta[i][row, :c] = torch.as_tensor(root.betas[j][p.name], dtype=f32).to(dev)
                        self.max_cnt[i] = max(self.max_cnt[i], c)
        self.stream = torch.cuda.current_stream(dev).cuda_stream

    # ---- DomainsList surface ---------------------------------------------------------------------
    def __len__(self):
        return self.n

    @property
    def minimum_lowers(self) -> float:
        """NS/heuristic/domains_list.py:315-321 (min over the stored output bounds; 1e-6 when empty)."""
        if self.n == 0:
            return 1e-6
        return float((self.lb[:self.n] - self.rhs[:self.n]).max(1).values.min().item())

    def _grow(self, need: int):
        if need <= self.cap:
            return
        new = max(2 * self.cap, need)

        def g(t):
            out = torch.zeros(new, *t.shape[1:], dtype=t.dtype, device=t.device)
            out[:self.n] = t[:self.n]
            return out
        for name in ('ids', 'lb', 'cs', 'rhs', 'x_L', 'x_U'):
            setattr(self, name, g(getattr(self, name)))
        for name in ('lower', 'upper', 'alpha', 'lA', 'h_cnt', 'h_loc', 'h_sign', 'h_beta'):
            setattr(self, name, [g(t) for t in getattr(self, name)])
        self.cap = new
        self.generation = getattr(self, 'generation', 0) + 1        # tensors moved: cached copy descriptors are stale

    def _grow_hist(self, k: int, J: int):
        if J <= self.Jc[k]:
            return
        newJ = max(2 * self.Jc[k], J)
        for name in ('h_loc', 'h_sign', 'h_beta'):
            t = getattr(self, name)[k]
            out = torch.zeros(self.cap, newJ, dtype=t.dtype, device=t.device)
            out[:, :self.Jc[k]] = t
            getattr(self, name)[k] = out
        self.Jc[k] = newJ
        self.generation = getattr(self, 'generation', 0) + 1

    def pick_out(self, batch: int) -> 'Picked':
        """The last `batch` records (the reference pops from the end of its storage), as views."""
        batch = min(batch, self.n)
        assert batch > 0
        self.visited += batch
        self.n -= batch
        return Picked(self, self.n, batch)


class Picked:
    """`batch` parent records at rows [row0, row0 + batch) of the store."""

    def __init__(self, store: DeviceDomainStore, row0: int, batch: int):
        self.store, self.row0, self.B = store, row0, batch

    def view(self, t: torch.Tensor) -> torch.Tensor:
        return t[self.row0:self.row0 + self.B]

    def results(self):
        """The picked parents as the reference's AbstractResults (device tensors), for callers of the host API."""
        from .abstractor import AbstractResults
        s = self.store
        lower = {p.name: self.view(s.lower[i]).view(self.B, *s.shape_k[i]) for i, p in enumerate(s.pres)}
        upper = {p.name: self.view(s.upper[i]).view(self.B, *s.shape_k[i]) for i, p in enumerate(s.pres)}
        lAs = {a.name: self.view(s.lA[i]).view(self.B, s.S, *s.shape_k[i]) for i, a in enumerate(s.acts)}
        slopes = {a.name: {s.final_name: self.view(s.alpha[i]).float().view(1, 1, self.B, -1).repeat(2, 1, 1, 1)}
                  for i, a in enumerate(s.acts)}
        hist, betas = [], []
        cnt = [self.view(c).cpu() for c in s.h_cnt]
        for b in range(self.B):
            h, bt = {}, {}
            for i, p in enumerate(s.pres):
                c = int(cnt[i][b])
                h[p.name] = (self.view(s.h_loc[i])[b, :c].long().cpu(), self.view(s.h_sign[i])[b, :c].cpu(), torch.zeros(c))
                bt[p.name] = self.view(s.h_beta[i])[b, :c].cpu()
            hist.append(h)
            betas.append(bt)
        masks = {k: ((lower[k] < 0) & (upper[k] > 0)).flatten(1).float() for k in lower}
        return AbstractResults(objective_ids=self.view(s.ids).cpu(), output_lbs=self.view(s.lb), masks=masks, lAs=lAs,
                               histories=hist, lower_bounds=lower, upper_bounds=upper,
                               input_lowers=self.view(s.x_L).view(self.B, *s.in_shape),
                               input_uppers=self.view(s.x_U).view(self.B, *s.in_shape), slopes=slopes, betas=betas,
                               cs=self.view(s.cs).view(self.B, s.S, s.n_out), rhs=self.view(s.rhs))


import contextlib
import os


@contextlib.contextmanager
def _trusted_indices(plan):
    """The split indices of the store were written by this library's own kernels: skip the per-call range check
    (one device reduction + sync) that capi runs on caller-supplied beta indices."""
    old = plan.validate_indices
    plan.validate_indices = False
    try:
        yield
    finally:
        plan.validate_indices = old


class DeviceBaB:
    """One hidden-split BaB iteration entirely on the device (Verifier._parallel_dpll steps 5-8,
    NS/verifier/verifier.py:373-405, without the host round trips)."""

    def __init__(self, net, store: DeviceDomainStore, decision_topk: int = 10, iteration: int = 20, lr_alpha: float = 0.1,
                 lr_beta: float = 0.1, lr_decay: float = 0.98, early_stop: bool = True, early_stop_patience: int = 10,
                 lookahead_rows: int = 1 << 17):
        self.net, self.store, self.plan = net, store, net.plan
        self.dev = store.dev
        self.topk = decision_topk
        self.opt = dict(iteration=iteration, lr_alpha=lr_alpha, lr_beta=lr_beta, lr_decay=lr_decay, early_stop=early_stop,
                        early_stop_patience=early_stop_patience)
        self.lookahead_rows = lookahead_rows
        self.offsets = [0]
        for nk in store.n_k:
            self.offsets.append(self.offsets[-1] + nk)
        self.n_total = self.offsets[-1]
        self.off_t = torch.tensor(self.offsets, dtype=torch.int64, device=self.dev)
        self.bias_vec = [self._bias_of(p) for p in store.pres]
        self.alpha_pos = []
        for a in store.acts:
            pos = None
            if a.alpha_indices is not None:
                idx = a.alpha_indices
                if isinstance(idx, (tuple, list)):
                    flat, stride = torch.zeros_like(idx[0]), 1
                    for d, ix in zip(reversed(a.output_shape[1:]), reversed(idx)):
                        flat = flat + ix * stride
                        stride *= int(d)
                    idx = flat
                n = 1
                for s_ in a.output_shape[1:]:
                    n *= int(s_)
                pos = capi.alpha_pos_from_index(idx, n, self.dev)
            self.alpha_pos.append(pos)
        self.stream = torch.cuda.current_stream(self.dev).cuda_stream
        self.last = {}
        self._child_cache = {}              # (rows, kind, history widths, store generation) -> child buffers + descriptor tables

    # ---- BaBSR bias term (heuristic/util.py:102-132) -----------------------------------------------
    def _bias_of(self, pre) -> Optional[torch.Tensor]:
        g = self.net._dev_graph()

        def conv_bias(nd, shape):
            b = nd.get('bias')
            if b is None:
                return None
            return b.view(-1, 1, 1).expand(shape).reshape(-1)

        nd = g[pre.index]
        shape = tuple(nd['shape'])
        if nd['op'] == 'linear':
            return None if nd.get('bias') is None else nd['bias'].contiguous()
        if nd['op'] == 'conv2d':
            b = conv_bias(nd, shape)
            return None if b is None else b.contiguous()
        if nd['op'] == 'batchnorm2d':
            return nd['bias'].view(-1, 1, 1).expand(shape).reshape(-1).contiguous()
        if nd['op'] == 'add':
            tot = torch.zeros(pre_numel(shape), device=self.dev)
            for j in nd['in']:
                sub = g[j]
                if sub['op'] == 'conv2d':
                    b = conv_bias(sub, shape)
                    if b is not None:
                        tot = tot + b
                elif sub['op'] == 'add':
                    for jj in sub['in']:
                        if g[jj]['op'] == 'conv2d':
                            b = conv_bias(g[jj], shape)
                            if b is not None:
                                tot = tot + b
            return tot.contiguous()
        raise NotImplementedError(f"BaBSR bias term of a pre-activation node produced by {nd['op']}")

    # ---- children ------------------------------------------------------------------------------------
    def _children(self, pick: Picked, src_rows: torch.Tensor, dec_layer, dec_neuron, side, with_history: bool, Jw=None):
        """Rows gathered from the store (src_rows: store row per child) with their split applied.
        Returns dict(C, x_L, x_U, rhs, lower, upper, alpha, beta or None, cnt)."""
        s, L = self.store, capi.lib()
        R = int(src_rows.numel())
        dev, f32 = self.dev, torch.float32
        # the child buffers and the copy-descriptor tables only depend on the row count, the history widths and where the
        # store's tensors live: a BaB loop asks for the same set every iteration, so they are built (and uploaded) once
        key = (R, bool(with_history), tuple(Jw) if Jw is not None else None, getattr(s, 'generation', 0))
        cache = self._child_cache
        hit = None if os.environ.get('CROWN_B200_NO_CHILD_CACHE') == '1' else cache.get(key)
        if hit is not None:
            ch, d_descs, n_descs, d_layers = hit
            ch = dict(ch)
        else:
            ch, d_descs, n_descs, d_layers = self._build_children(R, with_history, Jw)
            for k in [k for k in cache if k[1] == key[1]]:      # one set per kind (step / look-ahead): they are GBs
                del cache[k]
            cache[key] = (dict(ch), d_descs, n_descs, d_layers)
        capi._check(L.cb_store_multi_copy(d_descs.data_ptr(), n_descs, src_rows.data_ptr(), None, R, self.stream))
        capi._check(L.cb_store_apply_split(d_layers.data_ptr(), s.n_layers, dec_layer.data_ptr(), dec_neuron.data_ptr(),
                                           side.data_ptr(), None, R, self.stream))
        ch['_keep'] = (d_descs, d_layers, src_rows, dec_layer, dec_neuron, side)
        return ch

    def _build_children(self, R: int, with_history: bool, Jw):
        s = self.store
        dev, f32 = self.dev, torch.float32
        ch = {'C': torch.empty(R, s.S, s.n_out, dtype=f32, device=dev), 'rhs': torch.empty(R, s.S, dtype=f32, device=dev),
              'x_L': torch.empty(R, *s.in_shape, dtype=f32, device=dev), 'x_U': torch.empty(R, *s.in_shape, dtype=f32, device=dev),
              'lower': [torch.empty(R, *sh, dtype=f32, device=dev) for sh in s.shape_k],
              'upper': [torch.empty(R, *sh, dtype=f32, device=dev) for sh in s.shape_k],
              'alpha': [torch.empty(2, 1, R, na, dtype=f32, device=dev) for na in s.n_alpha]}
        descs = [_desc(s.cs, ch['C'], s.S * s.n_out, capi.COPY_F32), _desc(s.rhs, ch['rhs'], s.S, capi.COPY_F32),
                 _desc(s.x_L, ch['x_L'], s.n_in, capi.COPY_F32), _desc(s.x_U, ch['x_U'], s.n_in, capi.COPY_F32)]
        for i in range(s.n_layers):
            descs.append(_desc(s.lower[i], ch['lower'][i], s.n_k[i], capi.COPY_F32))
            descs.append(_desc(s.upper[i], ch['upper'][i], s.n_k[i], capi.COPY_F32))
            descs.append(_desc(s.alpha[i], ch['alpha'][i][0], s.n_alpha[i], capi.COPY_F16_TO_F32))
        layers = (CbSplitLayer * s.n_layers)()
        if with_history:
            ch['beta'], ch['cnt'] = [], []
            for i in range(s.n_layers):
                J = Jw[i]
                bt = {'val': torch.empty(R, J, dtype=f32, device=dev), 'loc': torch.empty(R, J, dtype=torch.int64, device=dev),
                      'sign': torch.empty(R, J, dtype=f32, device=dev), 'bias': None}
                cnt = torch.empty(R, dtype=torch.int32, device=dev)
                w = min(J, s.Jc[i])
                descs.append(_desc(s.h_beta[i], bt['val'], w, capi.COPY_F32, src_stride=s.Jc[i], dst_stride=J, dst_width=J))
                descs.append(_desc(s.h_sign[i], bt['sign'], w, capi.COPY_F32, src_stride=s.Jc[i], dst_stride=J, dst_width=J))
                descs.append(_desc(s.h_loc[i], bt['loc'], w, capi.COPY_I32_TO_I64, src_stride=s.Jc[i], dst_stride=J, dst_width=J))
                descs.append(_desc(s.h_cnt[i], cnt, 1, capi.COPY_I32))
                ch['beta'].append(bt)
                ch['cnt'].append(cnt)
        for i in range(s.n_layers):
            l = layers[i]
            l.lower, l.upper, l.n = ch['lower'][i].data_ptr(), ch['upper'][i].data_ptr(), s.n_k[i]
            if with_history:
                bt = ch['beta'][i]
                l.J = Jw[i]
                l.hist_cnt, l.hist_loc = ch['cnt'][i].data_ptr(), bt['loc'].data_ptr()
                l.hist_sign, l.beta_val, l.hist_point = bt['sign'].data_ptr(), bt['val'].data_ptr(), None
        d_descs = _descs_to_device(descs, dev)
        d_layers = torch.frombuffer(bytearray(bytes(layers)), dtype=torch.uint8).to(dev)
        if with_history:
            ch['cnt_tab'] = torch.tensor([c.data_ptr() for c in ch['cnt']], dtype=torch.int64, device=dev)
        return ch, d_descs, len(descs), d_layers

    # ---- branching (f2) --------------------------------------------------------------------------------
    def branch(self, pick: Picked):
        """BaBSR + top-k look-ahead (smart_hidden_branching): returns (layer [B] int32, neuron [B] int32) on the device."""
        s, L, B, dev = self.store, capi.lib(), pick.B, self.dev
        nt = self.n_total
        score = torch.empty(B, nt, device=dev)
        backup = torch.empty(B, nt, device=dev)
        mask = torch.empty(B, nt, device=dev)
        for i in range(s.n_layers):
            bias = self.bias_vec[i]
            capi._check(L.cb_babsr_scores(pick.view(s.lA[i]).data_ptr(), pick.view(s.lower[i]).data_ptr(),
                                          pick.view(s.upper[i]).data_ptr(), None if bias is None else bias.data_ptr(), B, s.S,
                                          s.n_k[i], score.data_ptr(), backup.data_ptr(), mask.data_ptr(), nt, self.offsets[i],
                                          self.stream))
        K = max(1, min(self.topk, nt))
        sv = torch.empty(B, K, device=dev)
        bv = torch.empty(B, K, device=dev)
        si = torch.empty(B, K, dtype=torch.int32, device=dev)
        bi = torch.empty(B, K, dtype=torch.int32, device=dev)
        capi._check(L.cb_topk_rows(score.data_ptr(), B, nt, K, 1, sv.data_ptr(), si.data_ptr(), self.stream))
        capi._check(L.cb_topk_rows(backup.data_ptr(), B, nt, K, 0, bv.data_ptr(), bi.data_ptr(), self.stream))
        # look-ahead: for every k the 2B candidates (score candidate of parent b, backup candidate of parent b), each
        # with its two children -> 4B rows, row = half * 2B + slot (decision_heuristics.py:90-157); all k in one pass
        lb_k = torch.empty(K, 4 * B, device=dev)
        parents = torch.arange(B, device=dev, dtype=torch.int32) + pick.row0
        kc = max(1, min(K, self.lookahead_rows // (4 * B)))
        rhs4 = pick.view(s.rhs).repeat(4, 1)
        for k0 in range(0, K, kc):
            k1 = min(K, k0 + kc)
            flat = torch.cat([si[:, k0:k1].t(), bi[:, k0:k1].t()], dim=1)                  # [kk, 2B] candidate per slot
            flat = torch.cat([flat, flat], dim=1).reshape(-1).long()                       # [kk * 4B]
            layer = (torch.bucketize(flat, self.off_t, right=True) - 1).to(torch.int32)
            neuron = (flat - self.off_t[layer.long()]).to(torch.int32)
            kk = k1 - k0
            src = parents.repeat(4 * kk)
            side = torch.cat([torch.ones(2 * B, device=dev), -torch.ones(2 * B, device=dev)]).repeat(kk)
            ch = self._children(pick, src, layer, neuron, side, with_history=False)
            lb, _ = self.plan.crown_pass(ch['C'], ch['x_L'], ch['x_U'], ch['lower'], ch['upper'], ch['alpha'], self.alpha_pos,
                                         None, want_lA=False)
            lb_k[k0:k1] = (lb - rhs4.repeat(kk, 1)).amax(dim=1).view(kk, 4 * B)
        dec_flat = torch.empty(B, dtype=torch.int32, device=dev)
        capi._check(L.cb_pick_decision(lb_k.data_ptr(), sv.data_ptr(), si.data_ptr(), bv.data_ptr(), bi.data_ptr(),
                                       mask.data_ptr(), nt, B, K, dec_flat.data_ptr(), self.stream))
        layer = (torch.bucketize(dec_flat.long(), self.off_t, right=True) - 1).to(torch.int32)
        neuron = (dec_flat.long() - self.off_t[layer.long()]).to(torch.int32)
        self.last['branch'] = dict(score=score, backup=backup, si=si, bi=bi, lb_k=lb_k)
        return layer, neuron

    # ---- one iteration -------------------------------------------------------------------------------------
    def step(self, batch: int, decisions=None) -> dict:
        """pick -> branch -> children -> alpha/beta-CROWN -> prune + append.  `decisions` (layer, neuron int32 device
        tensors) overrides the branching heuristic.  Returns counts (one small read-back)."""
        s, L, dev = self.store, capi.lib(), self.dev
        pick = s.pick_out(batch)
        B = pick.B
        layer, neuron = self.branch(pick) if decisions is None else decisions
        parents = torch.arange(B, device=dev, dtype=torch.int32) + pick.row0
        src = parents.repeat(2)
        side = torch.cat([torch.ones(B, device=dev), -torch.ones(B, device=dev)])
        Jw = [min(s.max_cnt[i], s.Jc[i]) + 1 for i in range(s.n_layers)]
        ch = self._children(pick, src, layer.repeat(2), neuron.repeat(2), side, with_history=True, Jw=Jw)
        o = self.opt
        with _trusted_indices(self.plan):
            lb, lA, n_iter = self.plan.optimize(ch['C'], ch['x_L'], ch['x_U'], ch['lower'], ch['upper'], ch['alpha'],
                                                self.alpha_pos, ch['beta'], ch['rhs'], iteration=o['iteration'],
                                                lr_alpha=o['lr_alpha'], lr_beta=o['lr_beta'], lr_decay=o['lr_decay'],
                                                early_stop_patience=o['early_stop_patience'], early_stop=o['early_stop'],
                                                want_lA=True)
        # prune lb > rhs, rank the survivors, append them after the remaining records
        R = 2 * B
        s._grow(s.n + R)
        rank = torch.empty(R, dtype=torch.int32, device=dev)
        out = torch.zeros(1 + s.n_layers, dtype=torch.int32, device=dev)
        cnt_tab = ch['cnt_tab']
        capi._check(L.cb_store_keep_rank(lb.data_ptr(), ch['rhs'].data_ptr(), R, s.S, s.n, rank.data_ptr(), out.data_ptr(),
                                         cnt_tab.data_ptr(), s.n_layers, self.stream))
        h_out = out.cpu()                                           # the iteration's one read-back
        kept = int(h_out[0])
        for i in range(s.n_layers):
            m = int(h_out[1 + i])
            s.max_cnt[i] = max(s.max_cnt[i], m)
            s._grow_hist(i, s.max_cnt[i] + 1)
        ids2 = pick.view(s.ids).repeat(2)
        descs = [_desc(ids2.view(torch.int32), s.ids.view(torch.int32), 2, capi.COPY_I32),
                 _desc(lb, s.lb, s.S, capi.COPY_F32), _desc(ch['C'], s.cs, s.S * s.n_out, capi.COPY_F32),
                 _desc(ch['rhs'], s.rhs, s.S, capi.COPY_F32), _desc(ch['x_L'], s.x_L, s.n_in, capi.COPY_F32),
                 _desc(ch['x_U'], s.x_U, s.n_in, capi.COPY_F32)]
        for i in range(s.n_layers):
            descs.append(_desc(ch['lower'][i], s.lower[i], s.n_k[i], capi.COPY_F32))
            descs.append(_desc(ch['upper'][i], s.upper[i], s.n_k[i], capi.COPY_F32))
            descs.append(_desc(ch['alpha'][i][0], s.alpha[i], s.n_alpha[i], capi.COPY_F32_TO_F16))
            descs.append(_desc(lA[i], s.lA[i], s.n_k[i], capi.COPY_F32, dst_stride=s.S * s.n_k[i], src_S=s.S, src_Bd=R))
            J, bt = Jw[i], ch['beta'][i]
            descs.append(_desc(bt['val'], s.h_beta[i], J, capi.COPY_F32, src_stride=J, dst_stride=s.Jc[i], dst_width=s.Jc[i]))
            descs.append(_desc(bt['sign'], s.h_sign[i], J, capi.COPY_F32, src_stride=J, dst_stride=s.Jc[i], dst_width=s.Jc[i]))
            descs.append(_desc(bt['loc'], s.h_loc[i], J, capi.COPY_I64_TO_I32, src_stride=J, dst_stride=s.Jc[i], dst_width=s.Jc[i]))
            descs.append(_desc(ch['cnt'][i], s.h_cnt[i], 1, capi.COPY_I32))
        d_descs = _descs_to_device(descs, dev)
        capi._check(L.cb_store_multi_copy(d_descs.data_ptr(), len(descs), None, rank.data_ptr(), R, self.stream))
        s.n += kept
        self.last.update(pick=pick, children=ch, lb=lb, lA=lA, rank=rank, layer=layer, neuron=neuron, n_iter=n_iter, keep=(d_descs, ids2))
        return {'picked': B, 'children': R, 'kept': kept, 'remaining': s.n, 'n_iter': n_iter}

    def run(self, batch: int, max_iterations: int = 1000) -> str:
        """'unsat' when no unverified domain is left (every leaf bounded above its threshold), else 'unknown'."""
        it = 0
        while len(self.store) > 0 and it < max_iterations:
            self.step(batch)
            it += 1
        self.iterations = it
        return 'unsat' if len(self.store) == 0 else 'unknown'


def pre_numel(shape) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n
