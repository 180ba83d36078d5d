"""ONNX -> torch.nn.Module for the operator set of the five BASELINE.json configs, WITHOUT the `onnx` package.

The reference loads networks with `onnx` + its vendored `onnx2pytorch` (NS/util/network/read_onnx.py:56-148,
NS/onnx2pytorch/convert/operations.py:57-429); neither is installable here, so this module reads the protobuf wire
format directly (ModelProto.graph = field 7; GraphProto.node = 1, .initializer = 5, .input = 11, .output = 12;
NodeProto.input = 1, .output = 2, .op_type = 4, .attribute = 5; TensorProto.dims = 1, .data_type = 2,
.float_data = 4, .int64_data = 7, .name = 8, .raw_data = 9) and rebuilds the network from nn.Linear / nn.Conv2d /
nn.BatchNorm2d / nn.ReLU / ... layers.  Same return convention as the reference's `parse_onnx`:
`(model, batched_input_shape, batched_output_shape, is_nhwc)`, BatchNormalization folded into the preceding Conv by
default (the reference's `merge_batch_norm` quirk, NS/onnx2pytorch/convert/operations.py:122-146).
"""
from __future__ import annotations

import struct
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn as nn


# ---- protobuf wire format -----------------------------------------------------------------------------------------
def _varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7


def _fields(buf: bytes):
    """Yield (field number, wire type, value) of one message; length-delimited values as memoryview slices."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            v = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            v = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f'unsupported protobuf wire type {wt}')
        yield fno, wt, v


def _signed(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [_signed(v)]
    out, pos, b = [], 0, bytes(v)
    while pos < len(b):
        x, pos = _varint(b, pos)
        out.append(_signed(x))
    return out


_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64}


def _tensor(buf: bytes):
    dims, dtype, name, raw = [], 1, '', None
    floats, int32s, int64s, doubles = [], [], [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            dims += _packed_varints(v, wt)
        elif fno == 2:
            dtype = v
        elif fno == 4:
            floats += list(struct.unpack(f'<{len(v) // 4}f', bytes(v))) if wt == 2 else [struct.unpack('<f', bytes(v))[0]]
        elif fno == 5:
            int32s += _packed_varints(v, wt)
        elif fno == 7:
            int64s += _packed_varints(v, wt)
        elif fno == 8:
            name = bytes(v).decode()
        elif fno == 9:
            raw = bytes(v)
        elif fno == 10:
            doubles += list(struct.unpack(f'<{len(v) // 8}d', bytes(v))) if wt == 2 else [struct.unpack('<d', bytes(v))[0]]
    np_dt = _DTYPES.get(dtype)
    if np_dt is None:
        raise NotImplementedError(f'ONNX tensor data_type {dtype}')
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt).copy()
    elif floats:
        arr = np.asarray(floats, dtype=np_dt)
    elif int64s:
        arr = np.asarray(int64s, dtype=np_dt)
    elif int32s:
        arr = np.asarray(int32s, dtype=np_dt)
    elif doubles:
        arr = np.asarray(doubles, dtype=np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    return name, arr.reshape(dims) if dims or arr.size == 1 else arr


def _attribute(buf: bytes):
    name, val = '', None
    ints, floats = [], []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:
            val = struct.unpack('<f', bytes(v))[0]
        elif fno == 3:
            val = _signed(v)
        elif fno == 4:
            val = bytes(v).decode(errors='replace')
        elif fno == 5:
            val = _tensor(bytes(v))[1]
        elif fno == 7:
            floats += list(struct.unpack(f'<{len(v) // 4}f', bytes(v))) if wt == 2 else [struct.unpack('<f', bytes(v))[0]]
        elif fno == 8:
            ints += _packed_varints(v, wt)
    if ints:
        val = ints
    elif floats:
        val = floats
    return name, val


def _node(buf: bytes) -> dict:
    nd = {'input': [], 'output': [], 'op': '', 'attr': {}, 'name': ''}
    for fno, wt, v in _fields(buf):
        if fno == 1:
            nd['input'].append(bytes(v).decode())
        elif fno == 2:
            nd['output'].append(bytes(v).decode())
        elif fno == 3:
            nd['name'] = bytes(v).decode()
        elif fno == 4:
            nd['op'] = bytes(v).decode()
        elif fno == 5:
            k, a = _attribute(bytes(v))
            nd['attr'][k] = a
    return nd


def _value_info(buf: bytes):
    name, shape = '', []
    for fno, wt, v in _fields(buf):
        if fno == 1:
            name = bytes(v).decode()
        elif fno == 2:                               # TypeProto
            for f2, _, v2 in _fields(bytes(v)):
                if f2 != 1:                          # tensor_type
                    continue
                for f3, _, v3 in _fields(bytes(v2)):
                    if f3 != 2:                      # shape
                        continue
                    for f4, _, v4 in _fields(bytes(v3)):
                        if f4 != 1:                  # dim
                            continue
                        dv = 0
                        for f5, w5, v5 in _fields(bytes(v4)):
                            if f5 == 1:
                                dv = _signed(v5)
                        shape.append(dv)
    return name, tuple(shape)


def load_onnx(path: str) -> dict:
    """{'nodes': [...], 'init': {name: ndarray}, 'inputs': [(name, shape)], 'outputs': [(name, shape)]}."""
    import gzip
    opener = gzip.open if path.endswith('.gz') else open
    with opener(path, 'rb') as f:
        buf = f.read()
    graph = None
    for fno, wt, v in _fields(buf):
        if fno == 7:
            graph = bytes(v)
    if graph is None:
        raise ValueError(f'{path}: no GraphProto')
    g = {'nodes': [], 'init': {}, 'inputs': [], 'outputs': []}
    for fno, wt, v in _fields(graph):
        if fno == 1:
            g['nodes'].append(_node(bytes(v)))
        elif fno == 5:
            name, arr = _tensor(bytes(v))
            g['init'][name] = arr
        elif fno == 11:
            g['inputs'].append(_value_info(bytes(v)))
        elif fno == 12:
            g['outputs'].append(_value_info(bytes(v)))
    g['inputs'] = [(n, s) for n, s in g['inputs'] if n not in g['init']]
    return g


# ---- graph -> nn.Module --------------------------------------------------------------------------------------------
class _Const(nn.Module):
    def __init__(self, value, sub=False, left=False):
        super().__init__()
        self.register_buffer('value', value)
        self.sub, self.left = sub, left

    def forward(self, x):
        if self.sub:
            return (self.value - x) if self.left else (x - self.value)
        return x + self.value


class OnnxModule(nn.Module):
    """Executes the converted layers in ONNX node order.  `steps` = [(kind, module name | None, input names, output
    name)]; every layer is a plain torch module so that `neuralsat_b200.graph.trace_module` (torch.fx) sees the
    operators of SURVEY.md section 8d."""

    def __init__(self, steps, layers: Dict[str, nn.Module], input_name: str, output_name: str, is_nhwc=False):
        super().__init__()
        self.steps = steps
        self.layers = nn.ModuleDict(layers)
        self.input_name, self.output_name = input_name, output_name
        self.is_nhwc = is_nhwc

    def forward(self, x):
        env = {self.input_name: x}
        for kind, mod, ins, out in self.steps:
            if kind == 'module':
                env[out] = self.layers[mod](env[ins[0]])
            elif kind == 'add':
                env[out] = env[ins[0]] + env[ins[1]]
            elif kind == 'sub':
                env[out] = env[ins[0]] - env[ins[1]]
            elif kind == 'flatten':
                env[out] = torch.flatten(env[ins[0]], 1)
            elif kind == 'identity':
                env[out] = env[ins[0]]
            else:
                raise NotImplementedError(kind)
        return env[self.output_name]


def add_batch(shape: tuple) -> tuple:
    """NS/util/network/read_onnx.py:43-51."""
    if len(shape) == 1:
        return (1, shape[0])
    if shape[0] not in (-1, 1):
        return (1, *shape)
    return tuple(shape)


def _pads(attr, k):
    p = attr.get('pads', [0] * (2 * k))
    if list(p[:k]) != list(p[k:]):
        raise NotImplementedError('asymmetric Conv padding')
    return tuple(int(v) for v in p[:k])


def convert(g: dict, merge_batch_norm: bool = True) -> OnnxModule:
    init = {k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in g['init'].items()}
    consts: Dict[str, torch.Tensor] = dict(init)          # names with a compile-time value
    steps, layers = [], {}
    producer = {}                                         # tensor name -> index in steps
    input_name = g['inputs'][0][0]
    n_consumers: Dict[str, int] = {}
    for nd in g['nodes']:
        for i in nd['input']:
            n_consumers[i] = n_consumers.get(i, 0) + 1

    def add_module(m, src, out):
        name = f'l{len(layers)}'
        layers[name] = m
        steps.append(('module', name, [src], out))
        producer[out] = len(steps) - 1

    for nd in g['nodes']:
        op, ins, out, at = nd['op'], nd['input'], nd['output'][0], nd['attr']
        cin = [consts.get(i) for i in ins]
        if op == 'Constant':
            consts[out] = torch.from_numpy(np.ascontiguousarray(at['value']))
        elif op in ('Shape', 'Gather', 'Unsqueeze', 'Concat', 'Cast') and all(c is not None for c in cin if True) and op != 'Shape':
            # shape arithmetic on constants feeding a Reshape: evaluate eagerly
            if op == 'Gather':
                consts[out] = cin[0][cin[1].long()] if cin[1].dim() else cin[0][int(cin[1])]
            elif op == 'Unsqueeze':
                axes = at.get('axes') or cin[1].tolist()
                t = cin[0]
                for a in sorted(axes):
                    t = t.unsqueeze(int(a))
                consts[out] = t
            elif op == 'Concat':
                consts[out] = torch.cat([c.reshape(-1) for c in cin], 0)
            else:
                consts[out] = cin[0]
        elif op == 'Shape':
            consts[out] = torch.tensor([-1], dtype=torch.int64)       # only ever used to rebuild "flatten" targets
        elif op == 'Gemm':
            W, b = cin[1], (cin[2] if len(ins) > 2 else None)
            if W is None or cin[0] is not None:
                raise NotImplementedError('Gemm with a non-constant weight')
            if at.get('transA', 0):
                raise NotImplementedError('Gemm transA')
            W = W.float() * float(at.get('alpha', 1.0))
            if not at.get('transB', 0):
                W = W.t()
            lin = nn.Linear(W.shape[1], W.shape[0], bias=b is not None)
            lin.weight.data = W.contiguous()
            if b is not None:
                lin.bias.data = (b.float() * float(at.get('beta', 1.0))).reshape(-1).contiguous()
            add_module(lin, ins[0], out)
        elif op == 'MatMul':
            if cin[1] is not None and cin[0] is None:           # x @ W
                W = cin[1].float().t().contiguous()
                lin = nn.Linear(W.shape[1], W.shape[0], bias=False)
                lin.weight.data = W
                add_module(lin, ins[0], out)
            elif cin[0] is not None and cin[1] is None:         # W @ x (column-vector convention, e.g. ACAS: [50,5] @ [5])
                W = cin[0].float().contiguous()
                lin = nn.Linear(W.shape[1], W.shape[0], bias=False)
                lin.weight.data = W
                add_module(lin, ins[1], out)
            else:
                raise NotImplementedError('MatMul of two activations')
        elif op in ('Add', 'Sub'):
            if cin[0] is None and cin[1] is None:
                steps.append(('add' if op == 'Add' else 'sub', None, [ins[0], ins[1]], out))
                producer[out] = len(steps) - 1
            elif cin[0] is not None and cin[1] is not None:
                consts[out] = cin[0] + cin[1] if op == 'Add' else cin[0] - cin[1]
            else:
                x, c, left = (ins[0], cin[1], False) if cin[0] is None else (ins[1], cin[0], True)
                c = c.float()
                # bias of the Linear that produced x: fold (MatMul + Add == Gemm), as onnx2pytorch does
                pi = producer.get(x)
                if (op == 'Add' and pi is not None and steps[pi][0] == 'module' and n_consumers.get(x, 0) == 1
                        and isinstance(layers[steps[pi][1]], nn.Linear) and layers[steps[pi][1]].bias is None
                        and c.numel() == layers[steps[pi][1]].out_features):
                    lin = layers[steps[pi][1]]
                    new = nn.Linear(lin.in_features, lin.out_features, bias=True)
                    new.weight.data = lin.weight.data
                    new.bias.data = c.reshape(-1).contiguous()
                    layers[steps[pi][1]] = new
                    steps[pi] = ('module', steps[pi][1], steps[pi][2], out)
                    producer[out] = pi
                else:
                    add_module(_Const(c, sub=(op == 'Sub'), left=left), x, out)
        elif op == 'Conv':
            W, b = cin[1].float(), (cin[2].float() if len(ins) > 2 else None)
            k = W.dim() - 2
            if k != 2:
                raise NotImplementedError('only Conv2d')
            conv = nn.Conv2d(W.shape[1] * int(at.get('group', 1)), W.shape[0], tuple(W.shape[2:]),
                             stride=tuple(at.get('strides', [1, 1])), padding=_pads(at, 2),
                             dilation=tuple(at.get('dilations', [1, 1])), groups=int(at.get('group', 1)),
                             bias=b is not None)
            conv.weight.data = W.contiguous()
            if b is not None:
                conv.bias.data = b.contiguous()
            add_module(conv, ins[0], out)
        elif op == 'BatchNormalization':
            gamma, beta, mean, var = (c.float() for c in cin[1:5])
            eps = float(at.get('epsilon', 1e-5))
            pi = producer.get(ins[0])
            prev = layers[steps[pi][1]] if pi is not None and steps[pi][0] == 'module' else None
            if merge_batch_norm and isinstance(prev, nn.Conv2d) and n_consumers.get(ins[0], 0) == 1:
                # NS/onnx2pytorch/convert/operations.py:122-146: fold into the preceding convolution
                scale = gamma / torch.sqrt(var + eps)
                new = nn.Conv2d(prev.in_channels, prev.out_channels, prev.kernel_size, prev.stride, prev.padding,
                                prev.dilation, prev.groups, bias=True)
                new.weight.data = (prev.weight.data * scale.view(-1, 1, 1, 1)).contiguous()
                pb = prev.bias.data if prev.bias is not None else torch.zeros_like(mean)
                new.bias.data = ((pb - mean) * scale + beta).contiguous()
                layers[steps[pi][1]] = new
                steps[pi] = ('module', steps[pi][1], steps[pi][2], out)
                producer[out] = pi
            else:
                bn = nn.BatchNorm2d(gamma.numel(), eps=eps)
                bn.weight.data, bn.bias.data = gamma.clone(), beta.clone()
                bn.running_mean.data, bn.running_var.data = mean.clone(), var.clone()
                add_module(bn.eval(), ins[0], out)
        elif op in ('Relu', 'Sigmoid', 'Tanh'):
            add_module({'Relu': nn.ReLU, 'Sigmoid': nn.Sigmoid, 'Tanh': nn.Tanh}[op](), ins[0], out)
        elif op == 'Flatten':
            steps.append(('flatten', None, [ins[0]], out))
            producer[out] = len(steps) - 1
        elif op == 'Reshape':
            # the configs only reshape to [batch, -1] (or to the same shape)
            steps.append(('flatten', None, [ins[0]], out))
            producer[out] = len(steps) - 1
        elif op in ('Identity', 'Dropout'):
            steps.append(('identity', None, [ins[0]], out))
            producer[out] = len(steps) - 1
        else:
            raise NotImplementedError(f'ONNX operator {op} is outside the hot-path operator set (SURVEY.md 8d)')
    out_name = g['outputs'][0][0]
    return OnnxModule(steps, layers, input_name, out_name).eval()


def parse_onnx(path: str, merge_batch_norm: bool = True):
    """Mirror of NS/util/network/read_onnx.py:parse_onnx: (model, batched_input_shape, batched_output_shape, is_nhwc)."""
    g = load_onnx(path)
    in_shape = tuple(d if d > 0 else 1 for d in g['inputs'][0][1])
    out_dims = g['outputs'][0][1]
    out_shape = tuple(d if d > 0 else 1 for d in out_dims) if len(out_dims) else (1,)
    model = convert(g, merge_batch_norm=merge_batch_norm)
    return model, add_batch(in_shape), add_batch(out_shape), False
