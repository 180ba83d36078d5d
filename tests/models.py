"""Small networks used by the golden generator and the parity tests.

Architectures are scaled-down instances of the BASELINE.json configs:
FC-ReLU (mnistfc), Conv-ReLU (oval21 base), Conv+BN+residual Add (sri_resnet_a / cifar100 resnets).
`toy_fixed` is the reference's fixed-weight ReLUNet (NS/example/test_model.py:80-108).
"""
import torch
import torch.nn as nn


class ToyFixed(nn.Module):
    def __init__(self):
        super().__init__()
        self.linear1 = nn.Linear(2, 3)
        self.linear2 = nn.Linear(3, 2)
        self.linear3 = nn.Linear(2, 3)
        with torch.no_grad():
            self.linear1.weight.copy_(torch.tensor([[1., 2.], [3., 4.], [5., 6.]]))
            self.linear1.bias.copy_(torch.tensor([1., 2., 3.]))
            self.linear2.weight.copy_(torch.tensor([[1., 2., 3.], [-4., -5., -6.]]))
            self.linear2.bias.copy_(torch.tensor([2., 3.]))
            self.linear3.weight.copy_(torch.tensor([[1., 2.], [-3., -4.], [-5., -6.]]))
            self.linear3.bias.copy_(torch.tensor([1., 2., 3.]))

    def forward(self, x):
        x = self.linear1(x).relu()
        x = self.linear2(x).relu()
        return self.linear3(x)


class ResBlock(nn.Module):
    """shortcut Conv1x1 s2 (+BN) || Conv3x3 s2 + BN + ReLU + Conv3x3 + BN  -> Add -> ReLU"""

    def __init__(self, cin, cout, bn=True):
        super().__init__()
        self.c1 = nn.Conv2d(cin, cout, 3, stride=2, padding=1)
        self.b1 = nn.BatchNorm2d(cout) if bn else nn.Identity()
        self.c2 = nn.Conv2d(cout, cout, 3, stride=1, padding=1)
        self.b2 = nn.BatchNorm2d(cout) if bn else nn.Identity()
        self.sc = nn.Conv2d(cin, cout, 1, stride=2)
        self.sb = nn.BatchNorm2d(cout) if bn else nn.Identity()
        self.r1 = nn.ReLU()
        self.r2 = nn.ReLU()

    def forward(self, x):
        y = self.b2(self.c2(self.r1(self.b1(self.c1(x)))))
        return self.r2(y + self.sb(self.sc(x)))


class ResNetSmall(nn.Module):
    def __init__(self, bn=True, cin=3, width=4, hw=8, n_out=5):
        super().__init__()
        self.c0 = nn.Conv2d(cin, width, 3, stride=1, padding=1)
        self.b0 = nn.BatchNorm2d(width) if bn else nn.Identity()
        self.r0 = nn.ReLU()
        self.blk = ResBlock(width, 2 * width, bn=bn)
        self.fc1 = nn.Linear(2 * width * (hw // 2) ** 2, 16)
        self.r = nn.ReLU()
        self.fc2 = nn.Linear(16, n_out)

    def forward(self, x):
        x = self.r0(self.b0(self.c0(x)))
        x = self.blk(x)
        x = torch.flatten(x, 1)
        return self.fc2(self.r(self.fc1(x)))


class NormalizedFC(nn.Module):
    """ACAS-Xu style front end: `x - mean` with a constant operand, then an FC ReLU net."""

    def __init__(self):
        super().__init__()
        self.register_buffer('mean', torch.linspace(0.1, 0.9, 10).view(1, 10))
        self.fc1 = nn.Linear(10, 24)
        self.fc2 = nn.Linear(24, 24)
        self.fc3 = nn.Linear(24, 4)

    def forward(self, x):
        x = x - self.mean
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))


def _randomize_bn(model, gen):
    """randomised running stats so BN is not the identity."""
    for m in model.modules():
        if isinstance(m, nn.BatchNorm2d):
            with torch.no_grad():
                m.running_mean.copy_(torch.rand(m.num_features, generator=gen) * 0.2 - 0.1)
                m.running_var.copy_(torch.rand(m.num_features, generator=gen) + 0.5)
                m.weight.copy_(torch.rand(m.num_features, generator=gen) + 0.5)
                m.bias.copy_(torch.rand(m.num_features, generator=gen) * 0.2 - 0.1)


def build_model(name, seed=0):
    """-> (nn.Module in eval mode, input shape without batch)."""
    torch.manual_seed(seed)
    if name == 'toy_fixed':
        return ToyFixed().eval(), (2,)
    if name == 'fc_small':
        m = nn.Sequential(nn.Linear(20, 32), nn.ReLU(), nn.Linear(32, 32), nn.ReLU(),
                          nn.Linear(32, 24), nn.ReLU(), nn.Linear(24, 5))
        return m.eval(), (20,)
    if name in ('fc_sigmoid', 'fc_tanh'):      # S-shaped activations (north star: sigmoid/tanh relaxations)
        act = nn.Sigmoid if name == 'fc_sigmoid' else nn.Tanh
        m = nn.Sequential(nn.Linear(12, 24), act(), nn.Linear(24, 20), act(), nn.Linear(20, 4))
        with torch.no_grad():       # wider pre-activations than default init so that all three cases occur
            for l in m:
                if isinstance(l, nn.Linear):
                    l.weight.mul_(3.0)
        return m.eval(), (12,)
    if name == 'fc_const':
        return NormalizedFC().eval(), (10,)
    if name == 'mnist_fc':          # BASELINE.json configs[1]
        m = nn.Sequential(nn.Flatten(), nn.Linear(784, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                          nn.Linear(256, 256), nn.ReLU(), nn.Linear(256, 256), nn.ReLU(),
                          nn.Linear(256, 10))
        return m.eval(), (1, 28, 28)
    if name == 'conv_small':        # oval21-base topology, scaled down
        m = nn.Sequential(nn.Conv2d(3, 4, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Conv2d(4, 8, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Flatten(), nn.Linear(8 * 2 * 2, 16), nn.ReLU(), nn.Linear(16, 5))
        return m.eval(), (3, 8, 8)
    if name == 'oval21_base':       # BASELINE.json configs[2] (cifar_base_kw topology)
        m = nn.Sequential(nn.Conv2d(3, 8, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Conv2d(8, 16, 4, stride=2, padding=1), nn.ReLU(),
                          nn.Flatten(), nn.Linear(16 * 8 * 8, 100), nn.ReLU(), nn.Linear(100, 10))
        return m.eval(), (3, 32, 32)
    if name == 'resnet_bn_small':
        m = ResNetSmall(bn=True)
        _randomize_bn(m, torch.Generator().manual_seed(seed + 1))
        return m.eval(), (3, 8, 8)
    if name == 'resnet_small':
        return ResNetSmall(bn=False).eval(), (3, 8, 8)
    raise KeyError(name)


# parameters of the reference BaB run that produces each fixture (oracle/gen_golden.py)
MODEL_SPECS = {
    'fc_small': dict(batch=8, n_iters=5, topk=2, eps=0.25, keep=(4, 3)),
    'mnist_fc': dict(batch=6, n_iters=3, topk=1, eps=0.03, keep=(1, 1)),
    'conv_small': dict(batch=6, n_iters=4, topk=2, eps=0.2, keep=(3, 2)),
    'resnet_bn_small': dict(batch=4, n_iters=3, topk=1, eps=0.1, keep=(2, 2)),
    'fc_const': dict(batch=6, n_iters=4, topk=2, eps=0.3, keep=(3, 2)),
    'fc_sigmoid': dict(batch=6, n_iters=4, topk=2, eps=0.5, keep=(3, 2)),
    'fc_tanh': dict(batch=6, n_iters=4, topk=2, eps=0.5, keep=(3, 2)),
    # BASELINE.json configs[2] at its real size (cifar_base_kw: 3x32x32 -> 8x16x16 -> 16x8x8 -> 100 -> 10)
    'oval21_base': dict(batch=4, n_iters=3, topk=1, eps=0.01, keep=(2, 1)),
}
