"""Graph kernels (csrc/graph/pagerank.cu): domain-level PageRank on the GPU (K13), with the NumPy oracle the CPU link
graph uses (infomesh_b200/index/link_graph.py; reference infomesh/index/link_graph.py:206-235)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from infomesh_b200 import _native

DAMPING = 0.85
MAX_ITERATIONS = 20
CONVERGENCE = 1e-6


def pagerank_ref(src, dst, weight, n, damping=DAMPING, max_iter=MAX_ITERATIONS, tol=CONVERGENCE):
    src, dst, w = np.asarray(src, np.int64), np.asarray(dst, np.int64), np.asarray(weight, np.float64)
    out_w = np.bincount(src, weights=w, minlength=n)
    share = np.divide(w, out_w[src], out=np.zeros_like(w), where=out_w[src] > 0)
    score = np.full(n, 1.0 / n)
    for _ in range(max_iter):
        nxt = np.full(n, (1.0 - damping) / n)
        np.add.at(nxt, dst, damping * score[src] * share)
        delta = float(np.abs(nxt - score).sum())
        score = nxt
        if delta < tol:
            break
    return score


def pagerank(src, dst, weight, n, damping=DAMPING, max_iter=MAX_ITERATIONS, tol=CONVERGENCE, device="cuda"):
    """Power iteration on the GPU; ``src`` / ``dst`` int edge endpoints, ``weight`` edge weights.  Returns fp32 ``[n]``."""
    dev = torch.device(device)
    s = torch.as_tensor(np.asarray(src), dtype=torch.int32, device=dev)
    d = torch.as_tensor(np.asarray(dst), dtype=torch.int32, device=dev)
    w = torch.as_tensor(np.asarray(weight), dtype=torch.float32, device=dev)
    out_w = torch.zeros(n, dtype=torch.float32, device=dev).index_add_(0, s.long(), w)
    share = torch.where(out_w[s.long()] > 0, w / out_w[s.long()].clamp(min=1e-30), torch.zeros_like(w)).contiguous()
    score = torch.full((n,), 1.0 / n, dtype=torch.float32, device=dev)
    delta = torch.zeros(1, dtype=torch.float32, device=dev)
    L = _native.require()
    for _ in range(max_iter):
        nxt = torch.full((n,), (1.0 - damping) / n, dtype=torch.float32, device=dev)
        delta.zero_()
        rc = L.im_pagerank_step(_native.ptr(s), _native.ptr(d), _native.ptr(share), ctypes.c_longlong(s.numel()),
                                ctypes.c_int(n), _native.ptr(score), _native.ptr(nxt), ctypes.c_float(damping),
                                _native.ptr(delta), _native.stream_ptr())
        _native.check(rc, "im_pagerank_step")
        _native.count_launch()
        score = nxt
        if float(delta.item()) < tol:
            break
    return score
