"""Epilogue cost study on the FFN-up / QKV shapes: same GEMM with different epilogues, CUDA-event timed."""
import sys

import torch

from infomesh_b200.ops.gemm import linear

dev = torch.device("cuda:0")
M = 163840
only = sys.argv[1] if len(sys.argv) > 1 else ""
for (n, k) in ((3072, 768), (2304, 768), (768, 3072), (768, 768)):
    a = (torch.randn(M, k, device=dev) * 0.5).bfloat16()
    w = (torch.randn(n, k, device=dev) * 0.05).bfloat16()
    b = torch.randn(n, device=dev)
    r = torch.randn(M, n, device=dev).bfloat16()
    out = torch.empty((M, n), device=dev, dtype=torch.bfloat16)
    variants = {"plain": dict(), "bias": dict(bias=b), "bias+gelu": dict(bias=b, act="gelu"), "bias+relu": dict(bias=b, act="relu"),
                "bias+res": dict(bias=b, residual=r), "bn128 bias+gelu": dict(bias=b, act="gelu", bn=128)}
    for name, kw in variants.items():
        if only and only not in name:
            continue
        for _ in range(3):
            linear(a, w, out=out, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            linear(a, w, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        print(f"M={M} N={n} K={k} {name:16s} {us:8.1f} us  {2 * M * n * k / us / 1e6:7.0f} TF/s")
