"""The tcgen05 implicit-GEMM convolution kernels (crown_conv_tc.cu) alone, against torch in float64: every convolution
geometry of the BASELINE.json architectures, both directions, ragged row counts, accumulate mode."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (Cin, Hin, Cout, K, stride, pad)
GEOMS = [
    (3, 32, 8, 4, 2, 1), (8, 16, 16, 4, 2, 1),                                      # oval21 base
    (3, 32, 32, 3, 1, 1), (32, 32, 32, 4, 2, 1), (32, 16, 128, 4, 2, 1),            # cifar10_2_255
    (3, 32, 16, 3, 2, 1), (16, 16, 32, 1, 2, 0), (16, 16, 32, 3, 2, 1), (32, 8, 32, 3, 1, 1),
    (64, 4, 128, 1, 2, 0), (64, 4, 128, 3, 2, 1), (128, 2, 128, 3, 1, 1),           # sri_resnet_a
    (3, 32, 64, 3, 2, 0), (64, 15, 128, 3, 2, 1), (64, 15, 128, 1, 2, 0), (128, 8, 128, 3, 1, 1),
    (128, 8, 128, 3, 2, 1), (128, 4, 128, 3, 1, 1),                                 # cifar100 resnet medium
    (5, 7, 6, 3, 1, 0), (4, 9, 200, 5, 2, 2), (130, 6, 20, 2, 2, 0),                # odd shapes, > 128 channels
]


@pytest.mark.parametrize('geom', GEOMS)
@pytest.mark.parametrize('rows', [1, 37])
def test_transpose_conv_matches_torch(geom, rows):
    from neuralsat_b200 import capi
    Cin, Hin, Cout, K, s, p = geom
    g = torch.Generator().manual_seed(sum(geom) + rows)
    Hout = (Hin + 2 * p - K) // s + 1
    W = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    X = torch.randn(rows, Cout, Hout, Hout, generator=g)
    op = Hin - ((Hout - 1) * s - 2 * p + K)
    ref = F.conv_transpose2d(X.double(), W.double(), None, stride=s, padding=p, output_padding=op)
    ref_b = torch.einsum('rchw,c->r', X.double(), b.double())
    Y, br = capi.conv_tc(X.cuda(), W.cuda(), b.cuda(), (Hin, Hin), s, p, 0)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert (Y.cpu().double() - ref).abs().max() <= 5e-6 * scale, (Y.cpu().double() - ref).abs().max() / scale
    assert (br.cpu().double() - ref_b).abs().max() <= 1e-5 * float(ref_b.abs().max().clamp(min=1.0))
    # accumulate on top of an existing tensor (residual fan-out, backward_bound.py:691-709)
    Y0 = torch.randn(rows, Cin, Hin, Hin, generator=g)
    Y2, _ = capi.conv_tc(X.cuda(), W.cuda(), None, (Hin, Hin), s, p, 0, Y=Y0.cuda().clone())
    assert (Y2.cpu().double() - (ref + Y0.double())).abs().max() <= 5e-6 * max(scale, 1.0)


@pytest.mark.parametrize('geom', GEOMS)
@pytest.mark.parametrize('rows', [1, 37])
def test_conv_matches_torch(geom, rows):
    from neuralsat_b200 import capi
    Cin, Hin, Cout, K, s, p = geom
    g = torch.Generator().manual_seed(sum(geom) + rows + 1)
    W = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    b = torch.randn(Cout, generator=g)
    X = torch.randn(rows, Cin, Hin, Hin, generator=g)
    ref = F.conv2d(X.double(), W.double(), b.double(), stride=s, padding=p)
    Y, _ = capi.conv_tc(X.cuda(), W.cuda(), b.cuda(), (Hin, Hin), s, p, 1)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert (Y.cpu().double() - ref).abs().max() <= 5e-6 * scale, (Y.cpu().double() - ref).abs().max() / scale
