"""Bring-up diagnostic (not a test): contraction error of the tcgen05 path vs torch fp32 matmul."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from neuralsat_b200 import capi
torch.backends.cuda.matmul.allow_tf32 = False
for (rows, N, K) in [(128, 16, 16), (128, 128, 64), (256, 256, 256), (512, 256, 784), (300, 100, 40), (512, 784, 256)]:
    g = torch.Generator().manual_seed(1)
    X = torch.randn(rows, K, generator=g).cuda()
    W = torch.randn(N, K, generator=g).cuda()
    Y = capi.tc_gemm(X, W, None)
    ref = X.double() @ W.double().t()
    scale = X.double().abs() @ W.double().abs().t()
    e_tc = ((Y.double() - ref).abs() / scale).max().item()
    e_32 = (((X @ W.t()).double() - ref).abs() / scale).max().item()
    print(f'rows={rows} N={N} K={K}: tcgen05 rel err {e_tc:.3e}   torch fp32 matmul rel err {e_32:.3e}', flush=True)
