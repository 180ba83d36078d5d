"""ctypes binding of the C-ABI in include/crown_b200.h (libcrown_b200.so).

This is the only module that touches the shared library.  There is NO fallback: if the library
is missing or a call fails, a RuntimeError is raised (allocation failures use the exact message
prefix the reference's OOM back-off matches, NS/util/misc/torch_cuda_memory.py:58-61).
PyTorch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional, Sequence

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libcrown_b200.so')

OPS = {'input': 0, 'linear': 1, 'conv2d': 2, 'batchnorm2d': 3, 'add': 4, 'sub': 5, 'flatten': 6,
       'relu': 7, 'sigmoid': 8, 'tanh': 9, 'addconst': 10}
CB_ERR_OOM = 3

c_float_p = C.POINTER(C.c_float)
c_i32_p = C.POINTER(C.c_int32)
c_i64_p = C.POINTER(C.c_int64)


class CbNode(C.Structure):
    _fields_ = [('op', C.c_int32), ('in0', C.c_int32), ('in1', C.c_int32),
                ('c', C.c_int32), ('h', C.c_int32), ('w', C.c_int32),
                ('weight', C.c_void_p), ('bias', C.c_void_p),
                ('kh', C.c_int32), ('kw', C.c_int32), ('stride_h', C.c_int32), ('stride_w', C.c_int32),
                ('pad_h', C.c_int32), ('pad_w', C.c_int32), ('dil_h', C.c_int32), ('dil_w', C.c_int32),
                ('groups', C.c_int32)]


class CbProblem(C.Structure):
    _fields_ = [('Bd', C.c_int32), ('S', C.c_int32),
                ('C', C.c_void_p), ('x_L', C.c_void_p), ('x_U', C.c_void_p),
                ('lower', C.POINTER(C.c_void_p)), ('upper', C.POINTER(C.c_void_p)),
                ('alpha', C.POINTER(C.c_void_p)), ('alpha_pos', C.POINTER(C.c_void_p)),
                ('n_alpha', c_i32_p), ('alpha_S1', C.c_int32),
                ('beta_val', C.POINTER(C.c_void_p)), ('beta_loc', C.POINTER(C.c_void_p)),
                ('beta_sign', C.POINTER(C.c_void_p)), ('beta_bias', C.POINTER(C.c_void_p)),
                ('beta_J', c_i32_p),
                ('lb', C.c_void_p), ('lA', C.POINTER(C.c_void_p))]


class CbOpt(C.Structure):
    _fields_ = [('iteration', C.c_int32), ('lr_alpha', C.c_float), ('lr_beta', C.c_float),
                ('lr_decay', C.c_float), ('early_stop_patience', C.c_int32),
                ('start_save_best', C.c_float), ('enable_beta', C.c_int32),
                ('early_stop', C.c_int32), ('rhs', C.c_void_p)]


class CbCopyDesc(C.Structure):
    _fields_ = [('src', C.c_void_p), ('dst', C.c_void_p), ('width', C.c_int32), ('dst_width', C.c_int32),
                ('mode', C.c_int32), ('src_S', C.c_int32), ('src_Bd', C.c_int32),
                ('src_stride', C.c_int64), ('dst_stride', C.c_int64)]


class CbSplitLayer(C.Structure):
    _fields_ = [('lower', C.c_void_p), ('upper', C.c_void_p), ('n', C.c_int32), ('J', C.c_int32),
                ('hist_cnt', C.c_void_p), ('hist_loc', C.c_void_p), ('hist_sign', C.c_void_p),
                ('beta_val', C.c_void_p), ('hist_point', C.c_void_p)]


COPY_F32, COPY_F16_TO_F32, COPY_F32_TO_F16, COPY_I32, COPY_I32_TO_I64, COPY_I64_TO_I32 = range(6)

_lib = None


def lib():
    """Load libcrown_b200.so (once).  Raises if it has not been built: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or neuralsat_b200/csrc/build.sh). neuralsat_b200 has no CPU fallback.')
    L = C.CDLL(LIB_PATH)
    L.cb_last_error.restype = C.c_char_p
    L.cb_version.restype = C.c_int
    L.cb_plan_create.argtypes = [C.POINTER(CbNode), C.c_int32, C.POINTER(C.c_void_p)]
    L.cb_plan_create.restype = C.c_int
    L.cb_plan_destroy.argtypes = [C.c_void_p]
    L.cb_plan_destroy.restype = None
    L.cb_plan_num_activations.argtypes = [C.c_void_p]
    L.cb_plan_num_activations.restype = C.c_int32
    L.cb_plan_activation_node.argtypes = [C.c_void_p, C.c_int32]
    L.cb_plan_activation_node.restype = C.c_int32
    L.cb_plan_preact_node.argtypes = [C.c_void_p, C.c_int32]
    L.cb_plan_preact_node.restype = C.c_int32
    L.cb_workspace_bytes.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(CbProblem)]
    L.cb_workspace_bytes.restype = C.c_size_t
    L.cb_crown_pass.argtypes = [C.c_void_p, C.POINTER(CbProblem), C.c_void_p, C.c_size_t, C.c_void_p]
    L.cb_crown_pass.restype = C.c_int
    L.cb_crown_grad.argtypes = [C.c_void_p, C.POINTER(CbProblem), C.POINTER(C.c_void_p),
                                C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_void_p]
    L.cb_crown_grad.restype = C.c_int
    L.cb_optimize.argtypes = [C.c_void_p, C.POINTER(CbProblem), C.POINTER(CbOpt), C.c_void_p,
                              C.c_size_t, C.c_void_p, c_i32_p]
    L.cb_optimize.restype = C.c_int
    L.cb_plan_uses_tensor_cores.argtypes = [C.c_void_p]
    L.cb_plan_uses_tensor_cores.restype = C.c_int32
    L.cb_plan_uses_chain.argtypes = [C.c_void_p]
    L.cb_plan_uses_chain.restype = C.c_int32
    L.cb_plan_uses_conv_tc.argtypes = [C.c_void_p]
    L.cb_plan_uses_conv_tc.restype = C.c_int32
    L.cb_plan_conv_choices.argtypes = [C.c_void_p, C.c_char_p, C.c_int32]
    L.cb_plan_conv_choices.restype = C.c_int32
    L.cb_debug_conv_tc.argtypes = [C.c_void_p] * 5 + [C.c_int32] * 11 + [C.c_void_p]
    L.cb_debug_conv_tc.restype = C.c_int
    L.cb_debug_conv_simt.argtypes = [C.c_void_p] * 4 + [C.c_int32] * 11 + [C.c_void_p]
    L.cb_debug_conv_simt.restype = C.c_int
    L.cb_debug_tc_gemm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                   C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
    L.cb_debug_tc_gemm.restype = C.c_int
    L.cb_debug_tc_times.argtypes = [C.c_void_p]
    L.cb_debug_tc_times.restype = None
    L.cb_store_multi_copy.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.cb_store_multi_copy.restype = C.c_int
    L.cb_store_apply_split.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.cb_store_apply_split.restype = C.c_int
    L.cb_store_keep_rank.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int32, C.c_void_p]
    L.cb_store_keep_rank.restype = C.c_int
    L.cb_babsr_scores.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
    L.cb_babsr_scores.restype = C.c_int
    L.cb_topk_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.cb_topk_rows.restype = C.c_int
    L.cb_pick_decision.argtypes = [C.c_void_p] * 6 + [C.c_int32] * 3 + [C.c_void_p, C.c_void_p]
    L.cb_pick_decision.restype = C.c_int
    L.cb_profile_enable.argtypes = [C.c_int32]
    L.cb_profile_enable.restype = None
    L.cb_launch_count.restype = C.c_int64
    L.cb_profile_num_kernels.restype = C.c_int32
    L.cb_profile_kernel_name.argtypes = [C.c_int32]
    L.cb_profile_kernel_name.restype = C.c_char_p
    L.cb_profile_collect.argtypes = [C.POINTER(C.c_double), c_i64_p, C.c_int32]
    L.cb_profile_collect.restype = C.c_int32
    _lib = L
    return L


def launch_count() -> int:
    return int(lib().cb_launch_count())


def profile_enable(on: bool):
    lib().cb_profile_enable(1 if on else 0)


def profile_collect() -> Dict[str, dict]:
    """{kernel class: {'ms': summed device time, 'launches': n}} since the last collect."""
    L = lib()
    n = L.cb_profile_num_kernels()
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    L.cb_profile_collect(ms, cnt, n)
    return {L.cb_profile_kernel_name(i).decode(): {'ms': ms[i], 'launches': int(cnt[i])}
            for i in range(n) if cnt[i] > 0}


EXPORTS = ['cb_last_error', 'cb_version', 'cb_plan_create', 'cb_plan_destroy',
           'cb_plan_num_activations', 'cb_plan_activation_node', 'cb_plan_preact_node',
           'cb_workspace_bytes', 'cb_crown_pass', 'cb_crown_grad', 'cb_optimize',
           'cb_plan_uses_tensor_cores', 'cb_plan_uses_chain', 'cb_plan_uses_conv_tc', 'cb_plan_conv_choices', 'cb_debug_conv_tc', 'cb_debug_conv_simt', 'cb_debug_tc_gemm', 'cb_debug_tc_times',
           'cb_store_multi_copy', 'cb_store_apply_split', 'cb_store_keep_rank', 'cb_babsr_scores', 'cb_topk_rows',
           'cb_pick_decision', 'cb_profile_enable', 'cb_launch_count', 'cb_profile_num_kernels',
           'cb_profile_kernel_name', 'cb_profile_collect']


def tc_gemm(X: torch.Tensor, W: torch.Tensor, col_bias: Optional[torch.Tensor] = None, bn: int = 0,
            dbg: int = 0) -> torch.Tensor:
    """Self-test hook: X[rows,K] @ W[N,K]^T (+ col_bias) on the tcgen05 3xTF32 kernel."""
    X, W = _f32(X, 'X'), _f32(W, 'W')
    rows, K = X.shape
    N = W.shape[0]
    Y = torch.empty(rows, N, dtype=torch.float32, device=X.device)
    stream = torch.cuda.current_stream(X.device).cuda_stream
    _check(lib().cb_debug_tc_gemm(X.data_ptr(), W.data_ptr(), _ptr(col_bias), Y.data_ptr(), rows, N, K, bn, dbg,
                                  stream))
    return Y


def conv_tc(X: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], in_hw, stride: int, pad: int, direction: int,
            Y: Optional[torch.Tensor] = None):
    """Self-test hook of the tensor-core convolution.  direction 0: conv_transpose2d(X, W) onto an in_hw map (+ the
    bias dot product, returned per row); direction 1: conv2d(X, W) + bias.  Returns (Y, bias_rows)."""
    X, W = _f32(X, 'X'), _f32(W, 'W')
    rows = int(X.shape[0])
    Cout, Cin, KH, KW = (int(v) for v in W.shape)
    Hin, Win = in_hw
    Hout, Wout = (Hin + 2 * pad - KH) // stride + 1, (Win + 2 * pad - KW) // stride + 1
    accumulate = Y is not None
    if Y is None:
        Y = torch.empty((rows, Cin, Hin, Win) if direction == 0 else (rows, Cout, Hout, Wout), dtype=torch.float32, device=X.device)
    br = torch.zeros(rows, dtype=torch.float32, device=X.device)
    stream = torch.cuda.current_stream(X.device).cuda_stream
    _check(lib().cb_debug_conv_tc(X.data_ptr(), W.data_ptr(), _ptr(bias), Y.data_ptr(), br.data_ptr(), rows, Cin, Hin, Win,
                                  Cout, KH, KW, stride, pad, direction, 1 if accumulate else 0, stream))
    return Y, br


def conv_simt(X: torch.Tensor, W: torch.Tensor, bias: Optional[torch.Tensor], in_hw, stride: int, pad: int,
              direction: int, Y: Optional[torch.Tensor] = None):
    """Self-test hook of the register-tiled fp32 convolution (thin first-layer kernels / tiled kernels); same
    conventions as conv_tc, returns Y."""
    X, W = _f32(X, 'X'), _f32(W, 'W')
    rows = int(X.shape[0])
    Cout, Cin, KH, KW = (int(v) for v in W.shape)
    Hin, Win = in_hw
    Hout, Wout = (Hin + 2 * pad - KH) // stride + 1, (Win + 2 * pad - KW) // stride + 1
    accumulate = Y is not None
    if Y is None:
        Y = torch.empty((rows, Cin, Hin, Win) if direction == 0 else (rows, Cout, Hout, Wout), dtype=torch.float32, device=X.device)
    stream = torch.cuda.current_stream(X.device).cuda_stream
    _check(lib().cb_debug_conv_simt(X.data_ptr(), W.data_ptr(), _ptr(bias), Y.data_ptr(), rows, Cin, Hin, Win, Cout, KH, KW,
                                    stride, pad, direction, 1 if accumulate else 0, stream))
    return Y


def _check(rc: int):
    if rc == 0:
        return
    msg = (lib().cb_last_error() or b'').decode()
    if rc == CB_ERR_OOM:
        raise RuntimeError('CUDA out of memory. ' + msg)
    raise RuntimeError(f'crown_b200 error {rc}: {msg}')


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _table(tensors: Sequence[Optional[torch.Tensor]]):
    arr = (C.c_void_p * max(1, len(tensors)))()
    for i, t in enumerate(tensors):
        arr[i] = _ptr(t)
    return arr


def _f32(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_cuda:
        raise TypeError(f'{name} must be a float32 CUDA tensor (got {t.dtype}, {t.device})')
    return t if t.is_contiguous() else t.contiguous()


def _shape3(shape):
    shape = tuple(int(s) for s in shape)
    if len(shape) == 3:
        return shape
    n = 1
    for s in shape:
        n *= s
    return (n, 1, 1)


class Plan:
    """Device-side plan of one network (cb_plan_t).  `nodes` is the list from graph.trace_module
    with all tensors already on the CUDA device."""

    def __init__(self, nodes: List[dict]):
        L = lib()
        self.nodes = nodes
        self._keep = []       # tensors the plan points into
        arr = (CbNode * len(nodes))()
        for i, nd in enumerate(nodes):
            cn = arr[i]
            cn.op = OPS[nd['op']]
            ins = nd.get('in', [])
            cn.in0 = ins[0] if len(ins) > 0 else -1
            cn.in1 = ins[1] if len(ins) > 1 else -1
            cn.c, cn.h, cn.w = _shape3(nd['shape'])
            cn.groups = 1
            if nd['op'] == 'linear':
                w = _f32(nd['weight'], 'weight')
                b = None if nd.get('bias') is None else _f32(nd['bias'], 'bias')
                self._keep += [w, b]
                cn.weight, cn.bias = _ptr(w), _ptr(b)
            elif nd['op'] == 'conv2d':
                w = _f32(nd['weight'], 'weight')
                b = None if nd.get('bias') is None else _f32(nd['bias'], 'bias')
                self._keep += [w, b]
                cn.weight, cn.bias = _ptr(w), _ptr(b)
                cn.kh, cn.kw = int(w.shape[2]), int(w.shape[3])
                cn.stride_h, cn.stride_w = nd['stride']
                cn.pad_h, cn.pad_w = nd['padding']
                cn.dil_h, cn.dil_w = nd['dilation']
                cn.groups = int(nd['groups'])
            elif nd['op'] == 'addconst':
                v = _f32(nd['value'], 'value')
                self._keep.append(v)
                cn.bias = _ptr(v)
            elif nd['op'] in ('sigmoid', 'tanh'):
                # tangent-point tables of the relaxation (auto_LiRPA/operators/tanh.py:65-130)
                from .sshape_tables import tangent_tables
                dev = next(t.device for t in self._keep if t is not None)
                d_lower, d_upper = tangent_tables(nd['op'], dev)
                self._keep += [d_lower, d_upper]
                cn.weight, cn.bias = _ptr(d_lower), _ptr(d_upper)
                cn.kh = int(d_lower.numel())
            elif nd['op'] == 'batchnorm2d':
                # folded affine form (auto_LiRPA/operators/normalization.py:117-118)
                scale = (nd['weight'] / torch.sqrt(nd['var'] + nd['eps'])).float().contiguous()
                shift = (nd['bias'] - nd['mean'] / torch.sqrt(nd['var'] + nd['eps']) * nd['weight']).float().contiguous()
                self._keep += [scale, shift]
                cn.weight, cn.bias = _ptr(scale), _ptr(shift)
        handle = C.c_void_p()
        _check(L.cb_plan_create(arr, len(nodes), C.byref(handle)))
        self.handle = handle
        self.n_act = L.cb_plan_num_activations(handle)
        self.tc_contractions = int(L.cb_plan_uses_tensor_cores(handle))
        self.conv_tc = int(L.cb_plan_uses_conv_tc(handle))
        buf = C.create_string_buffer(4096)
        self.conv_choices = buf.value.decode() if L.cb_plan_conv_choices(handle, buf, 4096) >= 0 else ''
        self.chain = bool(L.cb_plan_uses_chain(handle))
        self.chain_grad = int(L.cb_plan_uses_chain(handle)) == 2
        self.act_nodes = [L.cb_plan_activation_node(handle, k) for k in range(self.n_act)]
        self.pre_nodes = [L.cb_plan_preact_node(handle, k) for k in range(self.n_act)]
        self.act_numel = []
        for a in self.act_nodes:
            n = 1
            for s in nodes[a]['shape']:
                n *= int(s)
            self.act_numel.append(n)
        self._ws: Optional[torch.Tensor] = None
        self.device = next(t.device for t in self._keep if t is not None)

        def _numel(i):
            n = 1
            for d in nodes[i]['shape']:
                n *= int(d)
            return n
        self.n_in = _numel(next(i for i, nd in enumerate(nodes) if nd['op'] == 'input'))
        self.n_out = _numel(len(nodes) - 1)
        # beta 'loc' indexes shared memory and global arrays inside the kernels; the range check costs one device
        # reduction + sync per call.  Callers whose indices come from this library's own kernels (domain_store) and
        # timing loops switch it off.
        self.validate_indices = os.environ.get('CROWN_B200_VALIDATE_INDICES', '1') != '0'

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                lib().cb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # ------------------------------------------------------------------------------------
    def _workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            # torch raises torch.OutOfMemoryError (a RuntimeError with the reference's message)
            self._ws = torch.empty(int(nbytes * 1.05) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _problem(self, Cm, x_L, x_U, lower, upper, alpha, alpha_pos, beta, lb, lA):
        """Build cb_problem_t.  Per-activation lists are in activation order; entries may be None."""
        keep = []
        pr = CbProblem()
        if Cm.dim() != 3 or int(Cm.shape[2]) != self.n_out:
            raise ValueError(f'C must be [Bd, S, {self.n_out}], got {tuple(Cm.shape)}')
        Bd, S = int(Cm.shape[0]), int(Cm.shape[1])
        pr.Bd, pr.S = Bd, S
        Cm = _f32(Cm, 'C'); x_L = _f32(x_L, 'x_L'); x_U = _f32(x_U, 'x_U')
        if x_L.numel() != Bd * self.n_in or x_U.numel() != Bd * self.n_in:
            raise ValueError(f'x_L / x_U must hold Bd * {self.n_in} values')
        keep += [Cm, x_L, x_U]
        pr.C, pr.x_L, pr.x_U = _ptr(Cm), _ptr(x_L), _ptr(x_U)
        if len(lower) != self.n_act or len(upper) != self.n_act:
            raise ValueError(f'{self.n_act} intermediate bound tensors are required')
        lower = [_f32(t, 'lower') for t in lower]
        upper = [_f32(t, 'upper') for t in upper]
        for k, (l, u) in enumerate(zip(lower, upper)):
            if l.numel() != Bd * self.act_numel[k] or u.numel() != Bd * self.act_numel[k]:
                raise ValueError(f'intermediate bounds of activation {k} have the wrong size')
        keep += lower + upper
        tl, tu = _table(lower), _table(upper)
        keep += [tl, tu]
        pr.lower, pr.upper = tl, tu
        S1 = None

        def one_s1(v, k):
            nonlocal S1
            if v not in (1, S):
                raise ValueError(f'alpha of activation {k}: spec dimension {v} is neither 1 nor S = {S}')
            if S1 is not None and S1 != v:
                raise ValueError(f'alpha of activation {k}: spec dimension {v} differs from the other layers ({S1})')
            S1 = v

        if alpha is not None:
            if len(alpha) != self.n_act:
                raise ValueError(f'{self.n_act} alpha entries are required (None for a layer without slopes)')
            planes, n_alpha = [], (C.c_int32 * max(1, self.n_act))()
            pos = [None] * self.n_act if alpha_pos is None else list(alpha_pos)
            for k, a in enumerate(alpha):
                if a is None:
                    planes.append(None)
                    continue
                if a.dtype != torch.float32 or not a.is_cuda or not a.is_contiguous():
                    raise TypeError('alpha tensors must be contiguous float32 CUDA tensors')
                if self.nodes[self.act_nodes[k]]['op'] != 'relu':
                    # S-shapes: the reference's full [8,S1,Bd,*shape] tangent-point tensor (OP/tanh.py:54-63)
                    if a.dim() < 4 or a.shape[0] != 8 or a.shape[2] != Bd or a[0, 0, 0].numel() != self.act_numel[k]:
                        raise ValueError(f'alpha of S-shaped activation {k} must be [8,S1,Bd,*shape]')
                    one_s1(int(a.shape[1]), k)
                    n_alpha[k] = self.act_numel[k]
                    planes.append(a)
                    continue
                # the reference's [2,S1,Bd,*] tensor or directly plane 0 [S1,Bd,*]; '*' is the node shape (1 or 3
                # dims) or the flat list of kept neurons (1 dim), so the parity of the rank tells the two apart
                if a.dim() in (4, 6):
                    if a.shape[0] != 2:
                        raise ValueError(f'alpha of activation {k}: a rank-{a.dim()} tensor must be [2,S1,Bd,...]')
                    p0 = a[0]
                elif a.dim() in (3, 5):
                    p0 = a
                else:
                    raise ValueError(f'alpha of activation {k}: expected [2,S1,Bd,...] or [S1,Bd,...], got {tuple(a.shape)}')
                if int(p0.shape[1]) != Bd:
                    raise ValueError(f'alpha of activation {k}: batch dimension {int(p0.shape[1])} != Bd = {Bd}')
                one_s1(int(p0.shape[0]), k)
                n_alpha[k] = p0[0, 0].numel()
                if n_alpha[k] != self.act_numel[k] and (pos[k] is None):
                    raise ValueError(f'alpha of activation {k} holds {n_alpha[k]} of {self.act_numel[k]} neurons but no '
                                     'alpha_pos map was given')
                planes.append(p0)
            keep += planes
            ta = _table(planes)
            for k, pz in enumerate(pos):
                if pz is not None and (pz.dtype != torch.int32 or pz.numel() != self.act_numel[k] or not pz.is_cuda):
                    raise TypeError('alpha_pos must be an int32 CUDA tensor [n_k]')
            tp = _table(pos)
            keep += [ta, tp, n_alpha, pos]
            pr.alpha, pr.alpha_pos, pr.n_alpha = ta, tp, n_alpha
        pr.alpha_S1 = 1 if S1 is None else S1
        if beta is not None:
            if len(beta) != self.n_act:
                raise ValueError(f'{self.n_act} beta entries are required (None for a layer without splits)')
            vals, locs, signs, biases = [], [], [], []
            Js = (C.c_int32 * max(1, self.n_act))()
            bad = None
            for k, bt in enumerate(beta):
                if bt is None or bt['val'].shape[1] == 0:
                    vals.append(None); locs.append(None); signs.append(None); biases.append(None)
                    Js[k] = 0
                    continue
                v, lc = bt['val'], bt['loc']
                if v.dtype != torch.float32 or not v.is_contiguous() or lc.dtype != torch.int64:
                    raise TypeError('beta val must be contiguous float32 and loc int64')
                if not (v.is_cuda and lc.is_cuda and bt['sign'].is_cuda):
                    raise TypeError('beta val / loc / sign must be CUDA tensors')
                J = int(v.shape[1])
                if tuple(v.shape) != (Bd, J) or tuple(lc.shape) != (Bd, J) or tuple(bt['sign'].shape) != (Bd, J):
                    raise ValueError(f'beta of activation {k}: val, loc and sign must all be [Bd, J] = [{Bd}, {J}]')
                if bt.get('bias') is not None and tuple(bt['bias'].shape) != (Bd, J):
                    raise ValueError(f'beta of activation {k}: bias must be [Bd, J]')
                if self.validate_indices:
                    oob = ((lc < 0) | (lc >= self.act_numel[k])).any()
                    bad = oob if bad is None else (bad | oob)
                Js[k] = J
                vals.append(v); locs.append(lc.contiguous()); signs.append(_f32(bt['sign'], 'sign'))
                biases.append(None if bt.get('bias') is None else _f32(bt['bias'], 'bias'))
            if bad is not None and bool(bad.item()):
                raise ValueError('beta loc holds a neuron index outside its layer')
            tv, tlc, tsg, tbs = _table(vals), _table(locs), _table(signs), _table(biases)
            keep += [vals, locs, signs, biases, tv, tlc, tsg, tbs, Js]
            pr.beta_val, pr.beta_loc, pr.beta_sign, pr.beta_bias, pr.beta_J = tv, tlc, tsg, tbs, Js
        keep.append(lb)
        pr.lb = _ptr(lb)
        if lA is not None:
            tA = _table(lA)
            keep += [lA, tA]
            pr.lA = tA
        return pr, keep

    def _alloc_out(self, Bd, S, want_lA):
        lb = torch.empty(Bd, S, dtype=torch.float32, device=self.device)
        lA = None
        if want_lA:
            lA = [torch.empty(S, Bd, *self.nodes[a]['shape'], dtype=torch.float32, device=self.device)
                  for a in self.act_nodes]
        return lb, lA

    def crown_pass(self, Cm, x_L, x_U, lower, upper, alpha=None, alpha_pos=None, beta=None,
                   want_lA=True):
        """F1.  Returns (lb [Bd,S], lA list of [S,Bd,*shape] or None)."""
        L = lib()
        lb, lA = self._alloc_out(int(Cm.shape[0]), int(Cm.shape[1]), want_lA)
        pr, keep = self._problem(Cm, x_L, x_U, lower, upper, alpha, alpha_pos, beta, lb, lA)
        nbytes = L.cb_workspace_bytes(self.handle, pr.Bd, pr.S, 0, C.byref(pr))
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(L.cb_crown_pass(self.handle, C.byref(pr), ws.data_ptr(), ws.numel(), stream))
        return lb, lA

    def crown_grad(self, Cm, x_L, x_U, lower, upper, alpha, alpha_pos=None, beta=None):
        """d(sum lb)/d alpha[k] (plane-0 shape) and /d beta_val[k].  Returns (lb, lA, ga, gb)."""
        L = lib()
        lb, lA = self._alloc_out(int(Cm.shape[0]), int(Cm.shape[1]), True)
        pr, keep = self._problem(Cm, x_L, x_U, lower, upper, alpha, alpha_pos, beta, lb, lA)
        ga = []
        for k, a in enumerate(alpha):
            if a is None:
                ga.append(None)
                continue
            if self.nodes[self.act_nodes[k]]['op'] != 'relu':
                ga.append(torch.zeros_like(a))        # [8,S1,Bd,*]: planes 0,2,4,6 receive gradient
                continue
            p0 = a[0] if a.dim() >= 4 and a.shape[0] == 2 and a.shape[2] == pr.Bd else a
            ga.append(torch.zeros_like(p0))
        gb = [None] * self.n_act
        if beta is not None:
            gb = [None if (bt is None or bt['val'].shape[1] == 0) else torch.zeros_like(bt['val']) for bt in beta]
        tga, tgb = _table(ga), _table(gb)
        nbytes = L.cb_workspace_bytes(self.handle, pr.Bd, pr.S, 1, C.byref(pr))
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        _check(L.cb_crown_grad(self.handle, C.byref(pr), tga, tgb, ws.data_ptr(), ws.numel(), stream))
        return lb, lA, ga, gb

    def optimize(self, Cm, x_L, x_U, lower, upper, alpha, alpha_pos=None, beta=None, rhs=None,
                 iteration=20, lr_alpha=0.1, lr_beta=0.1, lr_decay=0.98, early_stop_patience=10,
                 start_save_best=0.5, enable_beta=True, early_stop=True, want_lA=True):
        """F2.  alpha[k] / beta[k]['val'] are updated IN PLACE to the reference's best snapshots.
        Returns (lb best [Bd,S], lA of the last pass, n_iter)."""
        L = lib()
        lb, lA = self._alloc_out(int(Cm.shape[0]), int(Cm.shape[1]), want_lA)
        pr, keep = self._problem(Cm, x_L, x_U, lower, upper, alpha, alpha_pos, beta, lb, lA)
        opt = CbOpt()
        opt.iteration = int(iteration)
        opt.lr_alpha, opt.lr_beta, opt.lr_decay = float(lr_alpha), float(lr_beta), float(lr_decay)
        opt.early_stop_patience = int(early_stop_patience)
        opt.start_save_best = float(start_save_best)
        opt.enable_beta = 1 if (enable_beta and beta is not None) else 0
        opt.early_stop = 1 if early_stop else 0
        if rhs is not None:
            rhs = _f32(rhs, 'rhs')
            if rhs.numel() != pr.Bd * pr.S:
                raise ValueError('rhs must be [Bd,S]')
        opt.rhs = _ptr(rhs)
        nbytes = L.cb_workspace_bytes(self.handle, pr.Bd, pr.S, 2, C.byref(pr))
        ws = self._workspace(nbytes)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        n_iter = C.c_int32(0)
        _check(L.cb_optimize(self.handle, C.byref(pr), C.byref(opt), ws.data_ptr(), ws.numel(), stream,
                             C.byref(n_iter)))
        return lb, lA, int(n_iter.value)


def alpha_pos_from_index(alpha_index: Optional[torch.Tensor], n: int, device) -> Optional[torch.Tensor]:
    """flattened neuron ids of the stored alphas (auto_LiRPA/operators/relu.py:330-332) ->
    int32 map neuron -> column (-1 = not stored)."""
    if alpha_index is None:
        return None
    pos = torch.full((n,), -1, dtype=torch.int32, device=device)
    idx = alpha_index.to(device=device, dtype=torch.long)
    pos[idx] = torch.arange(idx.numel(), dtype=torch.int32, device=device)
    return pos
