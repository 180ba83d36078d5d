"""tcgen05 3xTF32 contraction (crown_tc.cu) alone, through the C-ABI self-test entry: the result must
be fp32-faithful (the north star's 1e-5 relative on bounds needs ~1e-6 on each contraction)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

SHAPES = [  # rows, N, K, bn
    (128, 16, 16, 0), (128, 128, 16, 0), (256, 64, 32, 0), (200, 100, 40, 0), (1000, 256, 256, 0),
    (4096, 784, 256, 0), (4096, 256, 784, 0), (333, 10, 5, 0), (512, 256, 10, 0), (130, 24, 20, 0), (256, 784, 64, 0), (256, 96, 48, 96),
    (640, 256, 256, 128), (640, 256, 256, 32), (640, 250, 100, 0),
]


@pytest.mark.parametrize('rows,N,K,bn', SHAPES)
def test_tc_gemm_fp32_faithful(rows, N, K, bn):
    from neuralsat_b200 import capi
    g = torch.Generator().manual_seed(rows * 7 + N * 3 + K)
    X = torch.randn(rows, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    b = torch.randn(N, generator=g).cuda()
    Y = capi.tc_gemm(X, W, b, bn=bn)
    ref = X.double() @ W.double().t() + b.double()
    scale = (X.double().abs() @ W.double().abs().t()).clamp(min=1.0)
    err = ((Y.double() - ref).abs() / scale).max().item()
    # fp32 SIMT accumulation of the same product sits around 1e-7 on this scale
    assert err < 2e-6, err
