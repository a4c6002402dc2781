"""Dense retrieval kernels (csrc/search/sim_topk.cu): fused similarity-GEMM + top-k, and the
candidate-list merge (K3/K4 of SURVEY.md §2.4).

Reference behaviour being replaced: ``VectorStore.search`` -> ``collection.query`` (hnswlib ANN,
reference infomesh/index/vector_store.py:187-254) and the per-peer result merge
(reference infomesh/search/query.py:492-508).
"""
from __future__ import annotations

import ctypes

import torch

from infomesh_b200 import _native


def sim_topk_ref(q, docs, k, alive=None):
    """fp32 oracle: exact top-k of ``q @ docs.T`` (score desc, id asc on ties)."""
    scores = q.float() @ docs.float().t()
    if alive is not None:
        scores = scores.masked_fill(~alive.bool()[None, :], float("-inf"))
    k = min(k, docs.shape[0])
    vals, idx = torch.topk(scores, k, dim=1, largest=True, sorted=True)
    return vals, idx


def sim_topk_partials(q, docs, ktop=16, alive=None, max_ctas=0, thr_init=None):
    """Run the fused kernel; returns per-CTA candidate lists ``(scores[P,nq,ktop], ids[P,nq,ktop])``.

    ``thr_init``: fp32 ``[nq]`` view (any stride) of known lower bounds on each query's K-th best score; documents
    scoring below it are rejected by the threshold filter without touching the candidate lists."""
    assert q.is_cuda and docs.is_cuda and q.dtype == torch.bfloat16 and docs.dtype == torch.bfloat16
    assert q.stride(1) == 1 and docs.stride(1) == 1 and q.shape[1] == docs.shape[1]
    nq, dim = q.shape
    n_docs = docs.shape[0]
    L = _native.require()
    max_ctas = max_ctas or _native.sm_budget("scan")
    sms = L.im_sm_count()
    tiles = (n_docs + 127) // 128
    grid = max(1, min(tiles, sms if max_ctas <= 0 else min(sms, max_ctas)))
    # ktop == 0: threshold pre-pass, one (max score, slot id) per epilogue warp -> [grid * 4, nq, 1]
    lists, width = (grid * 4, 1) if ktop == 0 else (grid, ktop)
    out_s = torch.empty((lists, nq, width), device=q.device, dtype=torch.float32)
    out_i = torch.empty((lists, nq, width), device=q.device, dtype=torch.int32)
    if alive is not None:
        assert alive.dtype == torch.uint8 and alive.numel() >= n_docs
    rc = L.im_sim_topk(_native.ptr(q), _native.ptr(docs), ctypes.c_int(nq), ctypes.c_int(n_docs), ctypes.c_int(dim),
                       ctypes.c_int(q.stride(0)), ctypes.c_int(docs.stride(0)), ctypes.c_int(ktop),
                       _native.ptr(alive), _native.ptr(out_s), _native.ptr(out_i), ctypes.c_int(max_ctas),
                       _native.ptr(thr_init), ctypes.c_int(thr_init.stride(0) if thr_init is not None else 0),
                       _native.stream_ptr())
    if rc < 0:
        _native.check(rc, "im_sim_topk")
    assert rc == grid, (rc, grid)
    _native.count_launch()
    return out_s, out_i


def sim_topk_partials_f8(q8, q_scale, d8, d_scale, ktop=16, alive=None, max_ctas=0, thr_init=None):
    """fp8 twin of :func:`sim_topk_partials`: ``q8`` / ``d8`` uint8 e4m3 rows, ``q_scale`` / ``d_scale`` fp32 per row;
    scores in the lists are true dot products (``acc * d_scale * q_scale``)."""
    assert q8.is_cuda and d8.is_cuda and q8.dtype == torch.uint8 and d8.dtype == torch.uint8
    assert q8.stride(1) == 1 and d8.stride(1) == 1 and q8.shape[1] == d8.shape[1] and q8.shape[1] % 128 == 0
    assert q_scale.dtype == torch.float32 and d_scale.dtype == torch.float32 and d_scale.numel() >= d8.shape[0]
    nq, dim = q8.shape
    n_docs = d8.shape[0]
    L = _native.require()
    max_ctas = max_ctas or _native.sm_budget("scan")
    sms = L.im_sm_count()
    tiles = (n_docs + 127) // 128
    grid = max(1, min(tiles, sms if max_ctas <= 0 else min(sms, max_ctas)))
    lists, width = (grid * 4, 1) if ktop == 0 else (grid, ktop)
    out_s = torch.empty((lists, nq, width), device=q8.device, dtype=torch.float32)
    out_i = torch.empty((lists, nq, width), device=q8.device, dtype=torch.int32)
    rc = L.im_sim_topk_f8(_native.ptr(q8), _native.ptr(q_scale), _native.ptr(d8), _native.ptr(d_scale), ctypes.c_int(nq),
                          ctypes.c_int(n_docs), ctypes.c_int(dim), ctypes.c_int(q8.stride(0)), ctypes.c_int(d8.stride(0)),
                          ctypes.c_int(ktop), _native.ptr(alive), _native.ptr(out_s), _native.ptr(out_i), ctypes.c_int(max_ctas),
                          _native.ptr(thr_init), ctypes.c_int(thr_init.stride(0) if thr_init is not None else 0), _native.stream_ptr())
    if rc < 0:
        _native.check(rc, "im_sim_topk_f8")
    assert rc == grid, (rc, grid)
    _native.count_launch()
    return out_s, out_i


def rescore_topk(q, docs, cand_rows, k_out, id_offset=0, out_scores=None, out_ids=None):
    """Exact bf16 re-scoring of ``cand_rows`` int64 ``[nq, n <= 32]`` (local rows, -1 = empty) -> best ``k_out`` by
    true dot product (score desc, id asc), ids shifted by ``id_offset``."""
    nq, n = cand_rows.shape
    assert q.dtype == torch.bfloat16 and docs.dtype == torch.bfloat16 and cand_rows.dtype == torch.int64 and cand_rows.is_contiguous()
    if out_scores is None:
        out_scores = torch.empty((nq, k_out), device=q.device, dtype=torch.float32)
    if out_ids is None:
        out_ids = torch.empty((nq, k_out), device=q.device, dtype=torch.int64)
    L = _native.require()
    rc = L.im_rescore_topk(_native.ptr(q), ctypes.c_int(q.stride(0)), _native.ptr(docs), ctypes.c_int(docs.stride(0)),
                           ctypes.c_int(q.shape[1]), _native.ptr(cand_rows), ctypes.c_int(nq), ctypes.c_int(n), ctypes.c_int(k_out),
                           ctypes.c_int64(id_offset), _native.ptr(out_scores), _native.ptr(out_ids), _native.stream_ptr())
    _native.check(rc, "im_rescore_topk")
    _native.count_launch()
    return out_scores, out_ids


def sim_topk_f8(q8, q_scale, d8, d_scale, k=10, alive=None, id_offset=0, push=None, rescore=None, k_fetch=None):
    """Top-``k`` over an e4m3 shard.  ``rescore=(q_bf16, docs_bf16)``: fetch ``k_fetch`` (default ``min(32, 2k)``)
    candidates from the fp8 pass and re-rank them exactly against the bf16 rows, so quantisation noise can only cost a
    document that was outside the over-fetched list (measured recall: tests/test_gpu_kernels.py)."""
    kf = k if rescore is None else min(32, k_fetch or 2 * k)
    assert 1 <= k <= kf <= 32
    thr = sample_threshold_f8(q8, q_scale, d8, d_scale, kf, alive)
    ps, pi = sim_topk_partials_f8(q8, q_scale, d8, d_scale, ktop=kf, alive=alive, thr_init=thr)
    if rescore is None:
        return topk_merge(ps, pi, k, id_offset=id_offset, push=push)
    _s, rows = topk_merge(ps, pi, kf)
    rs, ri = rescore_topk(rescore[0], rescore[1], rows, k, id_offset=id_offset)
    if push is None:
        return rs, ri
    # the fused exchange lives in the merge kernel: a 1-list merge of the re-scored (already global) list pushes it
    return topk_merge(rs.unsqueeze(0), ri.unsqueeze(0), k, push=push)


def sample_threshold_f8(q8, q_scale, d8, d_scale, k, alive=None):
    """:func:`sample_threshold` for the fp8 shard."""
    n = d8.shape[0]
    n_s = max(SAMPLE_MIN_DOCS, n // SAMPLE_FRACTION)
    if n < 2 * n_s or k > 32:
        return None
    ps, pi = sim_topk_partials_f8(q8, q_scale, d8[:n_s], d_scale, ktop=0, alive=alive)
    if ps.shape[0] < k:
        return None
    ms, _ = topk_merge(ps, pi, k)
    return ms[:, k - 1]


def topk_merge(cand_scores, cand_ids, k_out, id_offset=0, out_scores=None, out_ids=None, push=None, wait=None):
    """Merge ``[P, nq, k_in]`` candidate lists into ``[nq, k_out]`` (score desc, id asc).

    ``push`` / ``wait`` are :class:`infomesh_b200.parallel.symm.TopkChannel` objects: with ``push`` the merged list is
    also stored into slot[rank] of every peer's receive area (fused exchange, no NCCL call); with ``wait`` the
    candidates are this rank's receive area and the kernel spins on the arrival counters first."""
    P, nq, k_in = cand_scores.shape
    assert cand_scores.dtype == torch.float32 and cand_ids.dtype in (torch.int32, torch.int64)
    cand_scores = cand_scores.contiguous()
    cand_ids = cand_ids.contiguous()
    if out_scores is None:
        out_scores = torch.empty((nq, k_out), device=cand_scores.device, dtype=torch.float32)
    if out_ids is None:
        out_ids = torch.empty((nq, k_out), device=cand_scores.device, dtype=torch.int64)
    L = _native.require()
    is64 = cand_ids.dtype == torch.int64
    rc = L.im_topk_merge(_native.ptr(cand_scores), _native.ptr(cand_ids if is64 else None),
                         _native.ptr(None if is64 else cand_ids), ctypes.c_int(P), ctypes.c_int(nq),
                         ctypes.c_int(k_in), ctypes.c_int(k_out), ctypes.c_int64(id_offset),
                         _native.ptr(out_scores), _native.ptr(out_ids),
                         ctypes.c_void_p(push.peer_scores_ptr if push else 0),
                         ctypes.c_void_p(push.peer_ids_ptr if push else 0),
                         ctypes.c_void_p(push.peer_flags_ptr if push else 0),
                         ctypes.c_int(push.world if push else 1), ctypes.c_int(push.rank if push else 0),
                         ctypes.c_void_p(wait.local_flags_ptr if wait else 0),
                         ctypes.c_void_p((push or wait).step_ptr if (push or wait) else 0), _native.stream_ptr(),
                         ctypes.c_void_p(wait.status_ptr if (wait is not None and wait.degraded_ok) else 0),
                         ctypes.c_uint(wait.wait_limit if wait is not None else 0))
    _native.check(rc, "im_topk_merge")
    _native.count_launch()
    return out_scores, out_ids


def sim_topk(q, docs, k=10, alive=None, id_offset=0, push=None):
    """Exact top-``k`` cosine/dot search of ``q[nq<=128, dim]`` against ``docs[n, dim]``.  With ``push`` the shard's
    result is also written into every peer's receive area by the merge kernel (fused top-k exchange)."""
    assert 1 <= k <= 32, "k > 32 is served by chunked search at the index level"
    ps, pi = sim_topk_partials(q, docs, ktop=k, alive=alive, thr_init=sample_threshold(q, docs, k, alive))
    return topk_merge(ps, pi, k, id_offset=id_offset, push=push)


SAMPLE_FRACTION = int(__import__("os").environ.get("INFOMESH_B200_SAMPLE_FRACTION", "32"))   # threshold sample = n_docs / 32 ...
SAMPLE_MIN_DOCS = 148 * 256   # ... but at least two 128-document tiles per SM; smaller shards skip the pre-pass


def sample_threshold(q, docs, k, alive=None):
    """Lower bound of every query's k-th best score, from a cheap pre-pass over the head of the shard.

    A persistent CTA starts with empty candidate lists, so without a bound the first tiles of every CTA insert almost
    every document (~k*ln(n/k) sorted inserts per query per warp -- a fixed ~1.9 ms at 64 queries that does not
    shrink with the shard).  The pre-pass runs the same kernel in max-only mode (``ktop=0``) over ``n/32`` documents
    (each epilogue warp reports the best score it saw per query, branch-free), and the k-th largest of those
    per-warp maxima -- k distinct documents --
    bounds the k-th best of the whole shard from below.  The main pass then inserts ~k*32 candidates per query
    per GPU in total.  Returns a strided fp32 ``[nq]`` view or ``None`` when the shard is too small to bother."""
    n = docs.shape[0]
    n_s = max(SAMPLE_MIN_DOCS, n // SAMPLE_FRACTION)
    if n < 2 * n_s or k > 32:
        return None
    ps, pi = sim_topk_partials(q, docs[:n_s], ktop=0, alive=alive)
    if ps.shape[0] < k:
        return None
    ms, _ = topk_merge(ps, pi, k)
    return ms[:, k - 1]
