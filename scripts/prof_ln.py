"""LayerNorm / embedding kernel timing at the reranker shape, CUDA events."""
import sys

import torch

from infomesh_b200.ops import nn as N

M = int(sys.argv[1]) if len(sys.argv) > 1 else 163840
dev = torch.device("cuda:0")
for H in (768, 384):
    x = torch.randn(M, H, device=dev).bfloat16()
    r = torch.randn(M, H, device=dev).bfloat16()
    g = torch.rand(H, device=dev) + 0.5
    b = torch.randn(H, device=dev)
    out = torch.empty_like(x)
    for name, fn, nbytes in (("sum_ln(x+res)", lambda: N.layernorm(x, g, b, 1e-12, residual=r, out=out), 3 * M * H * 2),
                             ("ln(x)", lambda: N.layernorm(x, g, b, 1e-12, out=out), 2 * M * H * 2)):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 50
        print(f"H={H} M={M} {name:14s} {us:7.1f} us  {nbytes / us / 1e6:.2f} TB/s")
