"""One GEMM shape/epilogue in a loop — ncu target.  usage: prof_gemm_one.py N K [bias|gelu|res|plain] [M]"""
import sys

import torch

from infomesh_b200.ops.gemm import linear

dev = torch.device("cuda:0")
n, k = int(sys.argv[1]), int(sys.argv[2])
var = sys.argv[3] if len(sys.argv) > 3 else "bias"
M = int(sys.argv[4]) if len(sys.argv) > 4 else 163840
a = (torch.randn(M, k, device=dev) * 0.5).bfloat16()
w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
b = torch.randn(n, device=dev)
r = torch.randn(M, n, device=dev).bfloat16()
out = torch.empty((M, n), device=dev, dtype=torch.bfloat16)
kw = {"plain": {}, "bias": dict(bias=b), "gelu": dict(bias=b, act="gelu"), "res": dict(bias=b, residual=r)}[var]
for _ in range(6):
    linear(a, w, out=out, **kw)
torch.cuda.synchronize()
