"""The register-tiled fp32 convolution kernels (crown_conv.cu) alone, against torch in float64: the thin first-layer
kernels (image side of at most four channels: every compiled-in (extent, stride, padding), odd map sizes, one-, three-
and four-channel images) and the tiled kernels they fall back to; both directions, ragged row counts, accumulate mode."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (Cin, Hin, Cout, K, stride, pad)
GEOMS = [
    (3, 32, 16, 3, 2, 1), (3, 32, 8, 4, 2, 1), (3, 32, 64, 3, 2, 0), (3, 32, 32, 3, 1, 1),       # the BASELINE first layers
    (3, 64, 64, 3, 2, 0), (1, 28, 16, 4, 2, 1), (1, 28, 32, 5, 2, 2), (1, 28, 8, 5, 1, 2),       # tinyimagenet, mnist convs
    (4, 9, 24, 3, 2, 1), (2, 11, 16, 4, 2, 0), (3, 7, 10, 3, 1, 1), (3, 13, 16, 5, 2, 2),        # odd sizes, ragged tiles
    (3, 10, 16, 2, 2, 0), (8, 16, 16, 4, 2, 1), (16, 8, 32, 3, 1, 1),                            # not thin: tiled kernels
]


@pytest.mark.parametrize('geom', GEOMS)
@pytest.mark.parametrize('rows', [1, 37])
def test_transpose_conv_matches_torch(geom, rows):
    from neuralsat_b200 import capi
    Cin, Hin, Cout, K, s, p = geom
    g = torch.Generator().manual_seed(sum(geom) + rows)
    Hout = (Hin + 2 * p - K) // s + 1
    W = torch.randn(Cout, Cin, K, K, generator=g) / (Cin * K * K) ** 0.5
    X = torch.randn(rows, Cout, Hout, Hout, generator=g)
    op = Hin - ((Hout - 1) * s - 2 * p + K)
    ref = F.conv_transpose2d(X.double(), W.double(), None, stride=s, padding=p, output_padding=op)
    Y = capi.conv_simt(X.cuda(), W.cuda(), None, (Hin, Hin), s, p, 0)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert (Y.cpu().double() - ref).abs().max() <= 2e-6 * scale, (Y.cpu().double() - ref).abs().max() / scale
    Y0 = torch.randn(rows, Cin, Hin, Hin, generator=g)
    Y2 = capi.conv_simt(X.cuda(), W.cuda(), None, (Hin, Hin), s, p, 0, Y=Y0.cuda().clone())
    assert (Y2.cpu().double() - (ref + Y0.double())).abs().max() <= 2e-6 * max(scale, 1.0)


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
    Y = capi.conv_simt(X.cuda(), W.cuda(), b.cuda(), (Hin, Hin), s, p, 1)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert (Y.cpu().double() - ref).abs().max() <= 2e-6 * scale, (Y.cpu().double() - ref).abs().max() / scale


def test_thin_kernels_are_the_ones_running(monkeypatch):
    """The first-layer geometries must not silently stay on the tiled kernels: with the thin kernels switched off the
    result is the same to rounding but not bit-identical (different summation order)."""
    import os
    from neuralsat_b200 import capi
    g = torch.Generator().manual_seed(5)
    W = torch.randn(16, 3, 3, 3, generator=g)
    X = torch.randn(5, 3, 32, 32, generator=g)
    Y1 = capi.conv_simt(X.cuda(), W.cuda(), None, (32, 32), 2, 1, 1)
    monkeypatch.setenv('CROWN_B200_DISABLE_CONV_THIN', '1')
    Y2 = capi.conv_simt(X.cuda(), W.cuda(), None, (32, 32), 2, 1, 1)
    assert torch.allclose(Y1, Y2, rtol=1e-5, atol=1e-5) and not torch.equal(Y1, Y2)
