"""Synthetic networks, specs and sub-domain batches for bench.py and the large-size tests.

Follows SURVEY.md section 8(d): random-init (torch.manual_seed) instances of the BASELINE.json
architectures, an L-inf box around x0~U(0,1), one margin row C = e_y - e_j, and a sub-domain batch
synthesised as in 8(d)(ii): every domain draws J~U{1..16} ReLU splits on unstable neurons with a
random sign and applies them to the intermediate bounds (child "active": l=0, child "inactive":
u=0, NS/abstractor/utils.py:233-247); alpha~U(0,1) rounded to fp16 (NS/abstractor/utils.py:51-59);
beta starts at 0 (auto_LiRPA/beta_crown.py:11-24).  Pure torch, runs on any device; the
intermediate bounds are plain interval bounds (sound, loose) — this is input fabrication, not
the measured path.
"""
from __future__ import annotations

from typing import Dict, List

import torch
import torch.nn as nn
import torch.nn.functional as F

from .graph import activation_indices, nodes_to, preact_indices, trace_module

WORKLOADS = {
    # BASELINE.json configs[1]: MNIST FC 256x4 ReLU, L-inf eps=0.02
    'mnistfc_256x4': dict(in_shape=(1, 28, 28), eps=0.02),
    # BASELINE.json configs[2]: CIFAR-10 oval21 base CNN
    'oval21_base': dict(in_shape=(3, 32, 32), eps=0.05),
    # BASELINE.json configs[0]: the ACAS Xu architecture (5 inputs, 6 x 50 ReLU, 5 outputs), here with hidden splits
    'acasxu': dict(in_shape=(5,), eps=0.05),
    # BASELINE.json configs[3]: CIFAR-10 ResNet, NB/sri_resnet_a (3 residual blocks 16-32-64-128, 1x1 stride-2
    # shortcuts, BN folded in the ONNX) and the plain CNN NB/cifar2020/cifar10_2_255_simplified
    'sri_resnet_a': dict(in_shape=(3, 32, 32), eps=0.02, bounds='center'),
    'cifar10_2_255': dict(in_shape=(3, 32, 32), eps=2.0 / 255, bounds='center'),
    # BASELINE.json configs[4]: NB/cifar100_tinyimagenet_resnet CIFAR100_resnet_medium / TinyImageNet_resnet_medium
    # (explicit BatchNormalization nodes, eps 1e-5, no ReLU after the residual Add)
    'cifar100_resnet_medium': dict(in_shape=(3, 32, 32), eps=0.02, bounds='center', fold_bn=True),
    'tinyimagenet_resnet_medium': dict(in_shape=(3, 56, 56), eps=0.02, bounds='center', fold_bn=True),
}


class _SriBlock(nn.Module):
    """shortcut Conv1x1 s2 || Conv3x3 s2 -> ReLU -> Conv3x3  -> Add -> ReLU   (sri_resnet_a, BN folded)"""

    def __init__(self, cin, cout):
        super().__init__()
        self.sc = nn.Conv2d(cin, cout, 1, stride=2)
        self.c1 = nn.Conv2d(cin, cout, 3, stride=2, padding=1)
        self.r1 = nn.ReLU()
        self.c2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1)
        self.r2 = nn.ReLU()

    def forward(self, x):
        return self.r2(self.sc(x) + self.c2(self.r1(self.c1(x))))


class _SriResNetA(nn.Module):
    def __init__(self):
        super().__init__()
        self.c0 = nn.Conv2d(3, 16, 3, stride=2, padding=1)
        self.r0 = nn.ReLU()
        self.b1, self.b2, self.b3 = _SriBlock(16, 32), _SriBlock(32, 64), _SriBlock(64, 128)
        self.fc1 = nn.Linear(512, 100)
        self.r = nn.ReLU()
        self.fc2 = nn.Linear(100, 10)

    def forward(self, x):
        x = self.b3(self.b2(self.b1(self.r0(self.c0(x)))))
        return self.fc2(self.r(self.fc1(torch.flatten(x, 1))))


def _cbn(cin, cout, k, s, p):
    return nn.Sequential(nn.Conv2d(cin, cout, k, stride=s, padding=p, bias=False), nn.BatchNorm2d(cout, eps=1e-5))


class _MedBlock(nn.Module):
    """Conv+BN+ReLU -> Conv+BN, plus the identity or a Conv1x1 s2 + BN shortcut; NO ReLU after the Add."""

    def __init__(self, cin, cout, stride, shortcut):
        super().__init__()
        self.a = _cbn(cin, cout, 3, stride, 1)
        self.r = nn.ReLU()
        self.b = _cbn(cout, cout, 3, 1, 1)
        self.sc = _cbn(cin, cout, 1, stride, 0) if shortcut else None

    def forward(self, x):
        y = self.b(self.r(self.a(x)))
        return y + (self.sc(x) if self.sc is not None else x)


class _ResNetMedium(nn.Module):
    """NB/cifar100_tinyimagenet_resnet/onnx/{CIFAR100,TinyImageNet}_resnet_medium.onnx: Conv 3->64 k3 s2 p0 + BN +
    ReLU, Conv 64->128 k3 s2 p1 + BN + ReLU, then 8 blocks at 128 channels (the first of each group of four has a
    Conv1x1+BN shortcut; the second group starts with stride 2), Flatten, Gemm -> n_hidden -> ReLU -> Gemm."""

    def __init__(self, hw, n_cls):
        super().__init__()
        self.stem = nn.Sequential(_cbn(3, 64, 3, 2, 0), nn.ReLU())
        h1 = (hw - 3) // 2 + 1                         # 15 (CIFAR) / 27 (TinyImageNet)
        # second stem conv output h2 x h2 feeds BOTH the first block's main path and its shortcut (from the stem output)
        self.pre = nn.ReLU()
        self.c2 = _cbn(64, 128, 3, 2, 1)
        h2 = (h1 + 2 - 3) // 2 + 1                     # 8 / 14
        self.c3 = _cbn(128, 128, 3, 1, 1)
        self.sc1 = _cbn(64, 128, 1, 2, 0)
        self.g1 = nn.ModuleList([_MedBlock(128, 128, 1, False) for _ in range(3)])
        self.g2a = _MedBlock(128, 128, 2, True)
        self.g2 = nn.ModuleList([_MedBlock(128, 128, 1, False) for _ in range(3)])
        h3 = (h2 + 2 - 3) // 2 + 1                     # 4 / 7
        self.fc1 = nn.Linear(128 * h3 * h3, n_cls)
        self.r = nn.ReLU()
        self.fc2 = nn.Linear(n_cls, n_cls)

    def forward(self, x):
        s = self.stem(x)                               # 64 x h1 x h1
        y = self.c3(self.pre(self.c2(s))) + self.sc1(s)
        for blk in self.g1:
            y = blk(y)
        y = self.g2a(y)
        for blk in self.g2:
            y = blk(y)
        return self.fc2(self.r(self.fc1(torch.flatten(y, 1))))


def build_nodes(name: str, seed: int = 0, fold_bn=None) -> List[dict]:
    """Node list of a workload.  BatchNorm is folded into the convolutions where the workload says so (what the
    reference's ONNX loader does with these files); fold_bn=False keeps the explicit BatchNorm nodes."""
    wl = WORKLOADS[name]
    fold = wl.get('fold_bn', False) if fold_bn is None else fold_bn
    return trace_module(build_network(name, seed), (1, *wl['in_shape']), fold_bn=fold)


def build_network(name: str, seed: int = 0) -> nn.Module:
    torch.manual_seed(seed)
    if name == 'mnistfc_256x4':
        m = nn.Sequential(nn.Flatten(), nn.Linear(784, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                          nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                          nn.Linear(256, 10))
    elif name == 'oval21_base':
        m = nn.Sequential(nn.Conv2d(3, 8, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Conv2d(8, 16, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Flatten(), nn.Linear(1024, 100), nn.ReLU(), nn.Linear(100, 10))
    elif name == 'acasxu':
        layers, w = [], 5
        for _ in range(6):
            layers += [nn.Linear(w, 50), nn.ReLU()]
            w = 50
        m = nn.Sequential(*layers, nn.Linear(50, 5))
    elif name == 'sri_resnet_a':
        m = _SriResNetA()
    elif name == 'cifar10_2_255':
        m = nn.Sequential(nn.Conv2d(3, 32, 3, stride=1, padding=1), nn.ReLU(),
                          nn.Conv2d(32, 32, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Conv2d(32, 128, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Flatten(), nn.Linear(8192, 250), nn.ReLU(), nn.Linear(250, 10))
    elif name in ('cifar100_resnet_medium', 'tinyimagenet_resnet_medium'):
        m = _ResNetMedium(32, 100) if name.startswith('cifar100') else _ResNetMedium(56, 200)
        # BN running statistics randomised as in SURVEY.md 8d (the default 0 / 1 would make BN a no-op)
        g = torch.Generator().manual_seed(seed + 1)
        for mod in m.modules():
            if isinstance(mod, nn.BatchNorm2d):
                c = mod.num_features
                mod.running_mean.data = torch.rand(c, generator=g) * 0.2 - 0.1
                mod.running_var.data = torch.rand(c, generator=g) + 0.5
                mod.weight.data = torch.rand(c, generator=g) + 0.5
                mod.bias.data = torch.rand(c, generator=g) * 0.2 - 0.1
    else:
        raise KeyError(name)
    return m.eval()


def _ibp(nodes: List[dict], x_L: torch.Tensor, x_U: torch.Tensor) -> Dict[int, tuple]:
    lo = [None] * len(nodes)
    hi = [None] * len(nodes)
    lo[0], hi[0] = x_L, x_U
    pre = {}
    for i, nd in enumerate(nodes):
        op = nd['op']
        if op == 'input':
            continue
        a_l, a_u = lo[nd['in'][0]], hi[nd['in'][0]]
        if op == 'linear':
            c, r = (a_l + a_u) / 2, (a_u - a_l) / 2
            cc, rr = F.linear(c, nd['weight'], nd.get('bias')), F.linear(r, nd['weight'].abs())
            lo[i], hi[i] = cc - rr, cc + rr
        elif op == 'conv2d':
            c, r = (a_l + a_u) / 2, (a_u - a_l) / 2
            args = (nd['stride'], nd['padding'], nd['dilation'], nd['groups'])
            cc = F.conv2d(c, nd['weight'], nd.get('bias'), *args)
            rr = F.conv2d(r, nd['weight'].abs(), None, *args)
            lo[i], hi[i] = cc - rr, cc + rr
        elif op == 'batchnorm2d':
            w = nd['weight'] / torch.sqrt(nd['var'] + nd['eps'])
            b = nd['bias'] - nd['mean'] * w
            c, r = (a_l + a_u) / 2, (a_u - a_l) / 2
            cc, rr = c * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1), r * w.abs().view(1, -1, 1, 1)
            lo[i], hi[i] = cc - rr, cc + rr
        elif op == 'add':
            lo[i], hi[i] = a_l + lo[nd['in'][1]], a_u + hi[nd['in'][1]]
        elif op == 'sub':
            lo[i], hi[i] = a_l - hi[nd['in'][1]], a_u - lo[nd['in'][1]]
        elif op == 'flatten':
            lo[i], hi[i] = a_l.flatten(1), a_u.flatten(1)
        elif op == 'relu':
            pre[nd['in'][0]] = (a_l, a_u)
            lo[i], hi[i] = F.relu(a_l), F.relu(a_u)
        else:
            raise NotImplementedError(op)
    return pre


def _center_bounds(nodes: List[dict], x0: torch.Tensor, rho: float) -> Dict[int, tuple]:
    """Fabricated intermediate bounds for deep networks, where interval bounds overflow: the concrete
    pre-activation values z at the box centre, +- rho * std(z) per layer (about a third of the neurons
    unstable).  Not sound; the arithmetic of the bounding path does not depend on soundness."""
    vals = [None] * len(nodes)
    vals[0] = x0
    pre = {}
    for i, nd in enumerate(nodes):
        op = nd['op']
        if op == 'input':
            continue
        a = vals[nd['in'][0]]
        if op == 'linear':
            vals[i] = F.linear(a, nd['weight'], nd.get('bias'))
        elif op == 'conv2d':
            vals[i] = F.conv2d(a, nd['weight'], nd.get('bias'), nd['stride'], nd['padding'], nd['dilation'], nd['groups'])
        elif op == 'batchnorm2d':
            w = nd['weight'] / torch.sqrt(nd['var'] + nd['eps'])
            b = nd['bias'] - nd['mean'] * w
            vals[i] = a * w.view(1, -1, 1, 1) + b.view(1, -1, 1, 1)
        elif op == 'add':
            vals[i] = a + vals[nd['in'][1]]
        elif op == 'sub':
            vals[i] = a - vals[nd['in'][1]]
        elif op == 'flatten':
            vals[i] = a.flatten(1)
        elif op == 'relu':
            r = rho * a.std().clamp(min=1e-6)
            pre[nd['in'][0]] = (a - r, a + r)
            vals[i] = F.relu(a)
        else:
            raise NotImplementedError(op)
    return pre


def make_batch(nodes: List[dict], Bd: int, eps: float, seed: int, device, max_splits: int = 16,
               bound_scale: float = 0.25, bounds: str = 'ibp'):
    """One batch of Bd sub-domains of the SAME root problem (same box, same C) as lists in
    activation order.  `bound_scale` shrinks the interval bounds of hidden layers towards their
    centre so that a realistic fraction of neurons is stable (pure IBP makes everything unstable);
    bounds='center' (the ResNet workloads) takes `_center_bounds` instead of interval bounds."""
    g = torch.Generator(device='cpu').manual_seed(seed)
    nodes = nodes_to(nodes, device)
    in_shape = tuple(nodes[0]['shape'])
    x0 = torch.rand(1, *in_shape, generator=g).to(device)
    x_L1, x_U1 = (x0 - eps).clamp(min=0), (x0 + eps).clamp(max=1)
    if bounds == 'center':
        pre = _center_bounds(nodes, (x_L1 + x_U1) / 2, 0.5)
        bound_scale = 1.0
    else:
        pre = _ibp(nodes, x_L1, x_U1)
    acts, pres = activation_indices(nodes), preact_indices(nodes)
    n_out = int(nodes[-1]['shape'][0])
    # one margin row e_0 - e_1 (S = 1, as in every classification config)
    C1 = torch.zeros(1, 1, n_out, device=device)
    C1[0, 0, 0] = 1.0
    C1[0, 0, 1 % n_out] = -1.0
    lower, upper, alpha, beta = [], [], [], []
    no_splits = max_splits <= 0              # un-split sub-domains (soundness checks): one dead beta slot per layer
    max_splits = max(max_splits, 1)
    layer_of_split = torch.randint(0, len(acts), (Bd, max_splits), generator=g)
    n_splits = torch.randint(1, max_splits + 1, (Bd,), generator=g)
    if no_splits:
        n_splits = torch.zeros(Bd, dtype=torch.int64)
    for k, p in enumerate(pres):
        l1, u1 = pre[p]
        c, r = (l1 + u1) / 2, (u1 - l1) / 2 * (bound_scale if k > 0 else 1.0)
        l = (c - r).expand(Bd, *l1.shape[1:]).clone()
        u = (c + r).expand(Bd, *u1.shape[1:]).clone()
        n = l[0].numel()
        lf, uf = l.view(Bd, n), u.view(Bd, n)
        unstable = ((lf[0] < 0) & (uf[0] > 0)).nonzero().flatten().cpu()
        J = max_splits
        loc = torch.zeros(Bd, J, dtype=torch.int64)
        sign = torch.zeros(Bd, J)
        if unstable.numel() > 0:
            pick = unstable[torch.randint(0, unstable.numel(), (Bd, J), generator=g)]
            sg = (torch.randint(0, 2, (Bd, J), generator=g) * 2 - 1).float()
            live = (layer_of_split == k) & (torch.arange(J).view(1, J) < n_splits.view(Bd, 1))
            # compact live entries to the front (SparseBeta layout, auto_LiRPA/beta_crown.py:25-42)
            order = torch.argsort((~live).to(torch.int8), dim=1, stable=True)
            live = torch.gather(live, 1, order)
            loc = torch.gather(pick, 1, order) * live
            sign = torch.gather(sg, 1, order) * live
            bi = torch.arange(Bd).view(Bd, 1).expand(Bd, J)
            act_m = (sign > 0)
            ina_m = (sign < 0)
            locd, bid = loc.to(device), bi.to(device)
            lf[bid[act_m.to(device)], locd[act_m.to(device)]] = 0.0
            uf[bid[ina_m.to(device)], locd[ina_m.to(device)]] = 0.0
        Jk = int((sign != 0).sum(1).max().item()) if Bd > 0 else 0
        Jk = max(Jk, 1)
        lower.append(l.contiguous())
        upper.append(u.contiguous())
        a = torch.rand(2, 1, Bd, *l.shape[1:], generator=g).half().float().to(device)
        alpha.append(a.contiguous())
        beta.append({'val': torch.zeros(Bd, Jk, device=device),
                     'loc': loc[:, :Jk].contiguous().to(device),
                     'sign': sign[:, :Jk].contiguous().to(device), 'bias': None})
    return {
        'C': C1.expand(Bd, 1, n_out).contiguous(),
        'x_L': x_L1.expand(Bd, *in_shape).contiguous(),
        'x_U': x_U1.expand(Bd, *in_shape).contiguous(),
        'lower': lower, 'upper': upper, 'alpha': alpha, 'beta': beta,
    }


def algorithmic_work(nodes: List[dict]):
    """Per-domain algorithmic work of ONE backward pass:
    MACs of the dense contractions, N_relu, N_in."""
    mac = 0
    n_relu = 0
    for nd in nodes:
        if nd['op'] == 'linear':
            mac += nd['weight'].numel()
        elif nd['op'] == 'conv2d':
            _, h, w = nd['shape']
            mac += nd['weight'].numel() * h * w
        elif nd['op'] in ('relu', 'sigmoid', 'tanh'):
            n = 1
            for s in nd['shape']:
                n *= int(s)
            n_relu += n
    n_in = 1
    for s in nodes[0]['shape']:
        n_in *= int(s)
    return {'mac': int(mac), 'n_relu': int(n_relu), 'n_in': int(n_in)}
