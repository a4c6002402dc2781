"""RRF fusion, cross-encoder pair assembly and final selection (csrc/search/fuse.cu) with CPU oracles."""
from __future__ import annotations

import ctypes

import torch

from infomesh_b200 import _native

RRF_K = 60.0


def rrf_fuse_ref(ids_a, ids_b, k_out, rrf_k=RRF_K, wa=1.0, wb=1.0):
    """Python oracle (reference infomesh/search/merge.py:37-133 semantics, keyed by doc id)."""
    out_s, out_i = [], []
    for ra, rb in zip(ids_a.tolist(), ids_b.tolist()):
        sc: dict[int, float] = {}
        for r, d in enumerate(ra):
            if d >= 0:
                sc[d] = sc.get(d, 0.0) + wa / (rrf_k + r + 1)
        for r, d in enumerate(rb):
            if d >= 0:
                sc[d] = sc.get(d, 0.0) + wb / (rrf_k + r + 1)
        order = sorted(sc.items(), key=lambda kv: (-kv[1], kv[0]))[:k_out]
        out_s.append([v for _, v in order] + [float("-inf")] * (k_out - len(order)))
        out_i.append([d for d, _ in order] + [-1] * (k_out - len(order)))
    return torch.tensor(out_s, dtype=torch.float32), torch.tensor(out_i, dtype=torch.int64)


def rrf_fuse(ids_a, ids_b, k_out, rrf_k=RRF_K, wa=1.0, wb=1.0, out_scores=None, out_ids=None, out_src=None):
    nq, ka = ids_a.shape
    kb = ids_b.shape[1]
    assert ids_a.dtype == torch.int64 and ids_b.dtype == torch.int64 and ids_a.is_contiguous() and ids_b.is_contiguous()
    if out_scores is None:
        out_scores = torch.empty((nq, k_out), device=ids_a.device, dtype=torch.float32)
    if out_ids is None:
        out_ids = torch.empty((nq, k_out), device=ids_a.device, dtype=torch.int64)
    L = _native.require()
    rc = L.im_rrf_fuse(_native.ptr(ids_a), _native.ptr(ids_b), ctypes.c_int(ka), ctypes.c_int(kb), ctypes.c_int(nq),
                       ctypes.c_float(rrf_k), ctypes.c_float(wa), ctypes.c_float(wb), ctypes.c_int(k_out),
                       _native.ptr(out_scores), _native.ptr(out_ids), _native.ptr(out_src), _native.stream_ptr())
    _native.check(rc, "im_rrf_fuse")
    _native.count_launch()
    return out_scores, out_ids


def ptr_table(tensors, device):
    """int64 device array of raw pointers (shard / peer tables)."""
    return torch.tensor([t.data_ptr() for t in tensors], dtype=torch.int64, device=device)


def build_pairs(q_tok, q_len, cand_ids, shard_tok_ptrs, shard_len_ptrs, docs_per_shard, passage_len, seq_len,
                bos=0, eos=2, pad=1, out_ids=None, out_lens=None):
    nq, max_q = q_tok.shape
    n_cand = cand_ids.shape[1]
    if out_ids is None:
        out_ids = torch.empty((nq * n_cand, seq_len), device=q_tok.device, dtype=torch.int32)
    if out_lens is None:
        out_lens = torch.empty((nq * n_cand,), device=q_tok.device, dtype=torch.int32)
    L = _native.require()
    rc = L.im_build_pairs(_native.ptr(q_tok), _native.ptr(q_len), ctypes.c_int(max_q), _native.ptr(cand_ids),
                          ctypes.c_int(nq), ctypes.c_int(n_cand), _native.ptr(shard_tok_ptrs),
                          _native.ptr(shard_len_ptrs), ctypes.c_longlong(docs_per_shard), ctypes.c_int(passage_len),
                          ctypes.c_int(seq_len), ctypes.c_int(bos), ctypes.c_int(eos), ctypes.c_int(pad),
                          _native.ptr(out_ids), _native.ptr(out_lens), _native.stream_ptr())
    _native.check(rc, "im_build_pairs")
    _native.count_launch()
    return out_ids, out_lens


def build_pairs_ref(q_tok, q_len, cand_ids, tok_store, len_store, passage_len, seq_len, bos=0, eos=2, pad=1):
    nq, max_q = q_tok.shape
    n_cand = cand_ids.shape[1]
    out = torch.full((nq * n_cand, seq_len), pad, dtype=torch.int32)
    lens = torch.ones((nq * n_cand,), dtype=torch.int32)
    for q in range(nq):
        ql = min(int(q_len[q]), max_q, seq_len // 2 - 2)
        for c in range(n_cand):
            d = int(cand_ids[q, c])
            p = q * n_cand + c
            pl = 0
            if d >= 0:
                pl = min(int(len_store[d]), passage_len, seq_len - ql - 4)
                pl = max(pl, 0)
            row = [bos] + q_tok[q, :ql].tolist() + [eos, eos] + (tok_store[d, :pl].tolist() if d >= 0 else []) + [eos]
            out[p, :len(row)] = torch.tensor(row, dtype=torch.int32)
            lens[p] = len(row) if d >= 0 else 1
    return out, lens


def rerank_select(logits, cand_ids, k_out, out_scores=None, out_ids=None):
    nq, n_cand = cand_ids.shape
    logits = logits.reshape(nq, n_cand)
    assert logits.dtype == torch.float32 and logits.is_contiguous() and cand_ids.is_contiguous()
    if out_scores is None:
        out_scores = torch.empty((nq, k_out), device=logits.device, dtype=torch.float32)
    if out_ids is None:
        out_ids = torch.empty((nq, k_out), device=logits.device, dtype=torch.int64)
    L = _native.require()
    rc = L.im_rerank_select(_native.ptr(logits), _native.ptr(cand_ids), ctypes.c_int(n_cand), ctypes.c_int(nq),
                            ctypes.c_int(k_out), _native.ptr(out_scores), _native.ptr(out_ids), _native.stream_ptr())
    _native.check(rc, "im_rerank_select")
    _native.count_launch()
    return out_scores, out_ids


_WEIGHT_CACHE: dict = {}


def rank_fuse(bm25, rows, k_out, *, crawled_at=None, trust=None, authority=None, title_match=None, url_path=None, now=None,
              weights=None, row_base: int = 0, want_signals: bool = False):
    """K12 on the device: the reference's six-signal ``combined_score`` + sort (infomesh/index/ranking.py:104-148,171-238).

    ``bm25`` fp32 / ``rows`` int64 ``[nq, n <= 32]`` are a retrieval list (rows < 0 = empty); ``crawled_at`` / ``trust`` /
    ``authority`` are per-DOCUMENT fp32 arrays indexed by ``row - row_base`` (``crawled_at`` holds seconds relative to the
    same epoch as ``now`` -- pass ages as ``now=0, crawled_at=-age`` to stay inside fp32); ``title_match`` / ``url_path`` are
    optional per-pair bonuses.  Returns (scores ``[nq, k_out]``, rows ``[nq, k_out]``[, signals ``[nq, k_out, 6]``])."""
    from infomesh_b200.index import ranking as R

    nq, n = rows.shape
    dev = rows.device
    key = (str(dev), tuple(sorted((weights or {}).items())))
    w = _WEIGHT_CACHE.get(key)
    if w is None:            # uploaded once per (device, override set): no H2D copy inside a captured step
        w = _WEIGHT_CACHE[key] = torch.tensor(R.weight_vector(weights), dtype=torch.float32).to(dev)
    out_s = torch.empty((nq, k_out), device=dev, dtype=torch.float32)
    out_r = torch.empty((nq, k_out), device=dev, dtype=torch.int64)
    sig = torch.zeros((nq, k_out, 6), device=dev, dtype=torch.float32) if want_signals else None
    L = _native.require()
    rc = L.im_rank_fuse(_native.ptr(bm25), _native.ptr(rows), ctypes.c_int(n), ctypes.c_int(nq), ctypes.c_longlong(row_base),
                        _native.ptr(crawled_at), _native.ptr(trust), _native.ptr(authority), _native.ptr(title_match),
                        _native.ptr(url_path), ctypes.c_float(0.0 if now is None else now),
                        ctypes.c_float(R.FRESHNESS_HALF_LIFE_SECONDS), ctypes.c_float(R.MIN_FRESHNESS), ctypes.c_float(R.DEFAULT_TRUST),
                        _native.ptr(w), ctypes.c_int(k_out), _native.ptr(out_s), _native.ptr(out_r), _native.ptr(sig), _native.stream_ptr())
    _native.check(rc, "im_rank_fuse")
    _native.count_launch()
    return (out_s, out_r, sig) if want_signals else (out_s, out_r)
