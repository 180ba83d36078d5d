"""nn.Module -> flat node list (the host-side "plan" source).

The reference obtains its graph by `torch.jit` tracing + ONNX-op mapping
(auto_LiRPA/bound_general.py:557-648, auto_LiRPA/parse_graph.py).  Here the graph comes from
`torch.fx.symbolic_trace`, restricted to the operators the five BASELINE.json configs contain
(SURVEY.md section 8d): Linear, Conv2d, BatchNorm2d, residual Add/Sub, Add/Sub of a constant, Flatten/Reshape,
ReLU, Sigmoid, Tanh.  Node order is program order == the order of the reference's `net.relus` /
`net.split_nodes`, which is what the golden fixtures rely on.

Each node is a dict (plain tensors, no nn.Module references):
  {'op','in':[idx...],'shape':tuple(without batch),'name':str, + op-specific tensors/attrs}
Node 0 is the input; the last node is the output (BoundedModule.final_name).
"""
from __future__ import annotations

import operator
from typing import List

import torch
import torch.fx as fx
import torch.nn as nn
import torch.nn.functional as F

_ACT_MODULES = {nn.ReLU: 'relu', nn.Sigmoid: 'sigmoid', nn.Tanh: 'tanh'}
_ACT_FUNCS = {F.relu: 'relu', torch.relu: 'relu', torch.sigmoid: 'sigmoid', F.sigmoid: 'sigmoid',
              torch.tanh: 'tanh', F.tanh: 'tanh'}
ACTIVATIONS = ('relu', 'sigmoid', 'tanh')


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (int(v), int(v))


class _Tracer(fx.Tracer):
    """Treat every supported layer as a leaf even when wrapped in custom containers."""

    def is_leaf_module(self, m, qualname):
        if isinstance(m, (nn.Linear, nn.Conv2d, nn.BatchNorm2d, nn.ReLU, nn.Sigmoid, nn.Tanh,
                          nn.Flatten, nn.Identity, nn.Dropout)):
            return True
        return False


def trace_module(model: nn.Module, input_shape, fold_bn: bool = False) -> List[dict]:
    """input_shape includes the batch dim (any value), like NetworkAbstractor.input_shape.
    fold_bn: merge every BatchNorm2d into the convolution in front of it (see `fold_batchnorm`)."""
    model = model.eval()
    graph = _Tracer().trace(model)
    gm = fx.GraphModule(model, graph)
    # shape propagation with a concrete forward
    from torch.fx.passes.shape_prop import ShapeProp
    dev = next((p.device for p in model.parameters()), torch.device('cpu'))
    ShapeProp(gm).propagate(torch.zeros(1, *tuple(input_shape)[1:], device=dev))

    nodes: List[dict] = []
    index = {}                      # fx node -> node index (after alias resolution)
    mods = dict(gm.named_modules())

    def shape_of(n):
        return tuple(n.meta['tensor_meta'].shape[1:])

    def add(n, d):
        d['shape'] = shape_of(n)
        d['name'] = f'/{len(nodes)}'
        nodes.append(d)
        index[n] = len(nodes) - 1

    def src(n):
        return index[n]

    consts = {}                     # fx get_attr node -> tensor (buffers / parameters used as operands)

    def const_of(a):
        if isinstance(a, fx.Node):
            return consts.get(a)
        if isinstance(a, (int, float)):
            return torch.tensor(float(a))
        if isinstance(a, torch.Tensor):
            return a
        return None

    for n in graph.nodes:
        if n.op == 'placeholder':
            if nodes:
                raise NotImplementedError('only single-input networks are supported')
            add(n, {'op': 'input', 'in': []})
        elif n.op == 'call_module':
            m = mods[n.target]
            if isinstance(m, nn.Linear):
                add(n, {'op': 'linear', 'in': [src(n.args[0])],
                        'weight': m.weight.detach().float().contiguous(),
                        'bias': None if m.bias is None else m.bias.detach().float().contiguous()})
            elif isinstance(m, nn.Conv2d):
                if m.padding_mode != 'zeros' or isinstance(m.padding, str):
                    raise NotImplementedError('only zero integer padding is supported')
                add(n, {'op': 'conv2d', 'in': [src(n.args[0])],
                        'weight': m.weight.detach().float().contiguous(),
                        'bias': None if m.bias is None else m.bias.detach().float().contiguous(),
                        'stride': _pair(m.stride), 'padding': _pair(m.padding),
                        'dilation': _pair(m.dilation), 'groups': int(m.groups)})
            elif isinstance(m, nn.BatchNorm2d):
                c = m.num_features
                w = m.weight.detach().float() if m.affine else torch.ones(c, device=m.running_mean.device)
                b = m.bias.detach().float() if m.affine else torch.zeros(c, device=m.running_mean.device)
                add(n, {'op': 'batchnorm2d', 'in': [src(n.args[0])], 'weight': w.contiguous(),
                        'bias': b.contiguous(), 'mean': m.running_mean.detach().float().contiguous(),
                        'var': m.running_var.detach().float().contiguous(), 'eps': float(m.eps)})
            elif type(m) in _ACT_MODULES:
                add(n, {'op': _ACT_MODULES[type(m)], 'in': [src(n.args[0])]})
            elif isinstance(m, nn.Flatten):
                self_in = src(n.args[0])
                if shape_of(n) == nodes[self_in]['shape']:
                    index[n] = self_in
                else:
                    add(n, {'op': 'flatten', 'in': [self_in]})
            elif isinstance(m, (nn.Identity, nn.Dropout)):
                index[n] = src(n.args[0])
            else:
                raise NotImplementedError(f'unsupported module {type(m).__name__}')
        elif n.op in ('call_function', 'call_method'):
            t = n.target
            if t in _ACT_FUNCS or t in ('relu', 'sigmoid', 'tanh'):
                add(n, {'op': _ACT_FUNCS.get(t, t), 'in': [src(n.args[0])]})
            elif t in (operator.add, torch.add, 'add', operator.iadd, operator.sub, torch.sub, 'sub'):
                is_sub = t in (operator.sub, torch.sub, 'sub')
                if n.kwargs.get('alpha', 1) != 1:
                    raise NotImplementedError('add/sub with a non-trivial alpha= scale')
                a, b = n.args[0], n.args[1]
                ca, cb_ = const_of(a), const_of(b)
                if ca is None and cb_ is None:
                    add(n, {'op': 'sub' if is_sub else 'add', 'in': [src(a), src(b)]})
                elif ca is None or not is_sub:
                    # x +/- constant (e.g. input normalisation `x - mean`): an unperturbed operand,
                    # auto_LiRPA/backward_bound.py:712-721 -> y = x + value, value has the node's shape
                    x_node, c = (a, cb_) if ca is None else (b, ca)
                    if const_of(x_node) is not None:
                        raise NotImplementedError('arithmetic between two constants')
                    value = (-c if is_sub else c).detach().float()
                    value = value.reshape(value.shape[1:]) if value.dim() == len(shape_of(n)) + 1 else value
                    value = value.expand(shape_of(n)).contiguous()
                    add(n, {'op': 'addconst', 'in': [src(x_node)], 'value': value})
                else:
                    raise NotImplementedError('constant - x')
            elif t in (torch.flatten, 'flatten', 'view', 'reshape', torch.reshape, 'contiguous', 'squeeze'):
                self_in = src(n.args[0])
                if t == 'contiguous' or shape_of(n) == nodes[self_in]['shape']:
                    index[n] = self_in
                else:
                    add(n, {'op': 'flatten', 'in': [self_in]})
            elif t in ('size', getattr) or t is getattr:
                index[n] = None       # shape arithmetic feeding view(); resolved by ShapeProp
            else:
                raise NotImplementedError(f'unsupported function {t}')
        elif n.op == 'get_attr':
            obj = gm
            for part in n.target.split('.'):
                obj = getattr(obj, part)
            consts[n] = torch.as_tensor(obj)
        elif n.op == 'output':
            out = n.args[0]
            if index[out] != len(nodes) - 1:
                raise NotImplementedError('the output must be the last computed node')
    for nd in nodes:
        if nd['op'] in ACTIVATIONS and nodes[nd['in'][0]]['op'] in ACTIVATIONS + ('input',):
            raise NotImplementedError('activation directly on an activation / the input')
    return fold_batchnorm(nodes) if fold_bn else nodes


def fold_batchnorm(nodes: List[dict]) -> List[dict]:
    """Conv2d -> BatchNorm2d (eval mode, the convolution's only consumer) becomes one Conv2d, as the reference's ONNX
    loader does by default (`merge_batch_norm`, NS/onnx2pytorch/convert/operations.py:122-146):
    W' = W * s[co], b' = (b - mean) * s + beta with s = gamma / sqrt(var + eps).  Node names are re-issued."""
    consumers = [0] * len(nodes)
    for nd in nodes:
        for j in nd.get('in', []):
            consumers[j] += 1
    remap, out = {}, []
    for i, nd in enumerate(nodes):
        if nd['op'] == 'batchnorm2d':
            j = nd['in'][0]
            src = nodes[j]
            if src['op'] == 'conv2d' and consumers[j] == 1:
                s = nd['weight'] / torch.sqrt(nd['var'] + nd['eps'])
                conv = out[remap[j]]
                b = conv.get('bias')
                b = torch.zeros_like(nd['mean']) if b is None else b
                conv['weight'] = (conv['weight'] * s.view(-1, 1, 1, 1)).contiguous()
                conv['bias'] = ((b - nd['mean']) * s + nd['bias']).contiguous()
                remap[i] = remap[j]
                continue
        new = dict(nd)
        new['in'] = [remap[j] for j in nd.get('in', [])]
        new['name'] = f'/{len(out)}'
        remap[i] = len(out)
        out.append(new)
    return out


def activation_indices(nodes: List[dict]) -> List[int]:
    """Indices of activation nodes in program order (== reference `net.relus` order)."""
    return [i for i, nd in enumerate(nodes) if nd['op'] in ACTIVATIONS]


def preact_indices(nodes: List[dict]) -> List[int]:
    """Indices of pre-activation nodes (== reference `net.split_nodes` order)."""
    return [nodes[i]['in'][0] for i in activation_indices(nodes)]


def nodes_to(nodes: List[dict], device) -> List[dict]:
    out = []
    for nd in nodes:
        out.append({k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in nd.items()})
    return out
