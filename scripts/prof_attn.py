"""Attention kernel timing at the reranker shape (B x 12 heads x 128 x 64), CUDA events."""
import sys

import torch

from infomesh_b200.ops.attention import attention

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1280
dev = torch.device("cuda:0")
qkv = (torch.randn(B, 128, 3 * 768, device=dev) * 0.7).bfloat16()
lens = torch.randint(100, 129, (B,), device=dev, dtype=torch.int32)
q, k, v = qkv[..., :768], qkv[..., 768:1536], qkv[..., 1536:]
for name, kl in (("full", None), ("kv_lens", lens)):
    for _ in range(3):
        attention(q, k, v, 12, kl)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        attention(q, k, v, 12, kl)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 50
    gb = (B * 128 * 768 * 2 * 4) / 1e9
    print(f"attention B={B} S=128 heads=12 hd=64 {name}: {us:.1f} us  ({gb / us * 1e6 / 1e3:.2f} TB/s of q+k+v+o)")
