"""BM25 inverted index on GPU (csrc/search/bm25.cu) + CSR builder (csrc/host/textproc.cpp) — K2/K11.

``Bm25Index`` owns the HBM-resident CSR postings of one shard.  ``bm25_ref`` is the NumPy oracle with
FTS5's formula (reference infomesh/index/local_store.py:316-340: ``bm25()`` defaults k1=1.2, b=0.75,
score reported as ``abs``).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np

from infomesh_b200 import _native

K1 = 1.2
B = 0.75
MAX_TERMS = 16


def idf_fts5(n_docs: int, df):
    """FTS5 idf: ln((N - n + 0.5) / (n + 0.5)), floored at 1e-6."""
    df = np.asarray(df, dtype=np.float64)
    v = np.log((n_docs - df + 0.5) / (df + 0.5))
    return np.maximum(v, 1e-6).astype(np.float32)


class HostIndexBuilder:
    """C++ tokeniser + posting builder (``IndexBuilder`` in csrc/host/textproc.cpp)."""

    def __init__(self):
        self.L = _native.lib()
        self.L.im_ib_create.restype = ctypes.c_void_p
        self.L.im_ib_tokenize.restype = ctypes.c_longlong
        self.L.im_ib_vocab.restype = ctypes.c_longlong
        self.L.im_ib_docs.restype = ctypes.c_longlong
        self.L.im_ib_nnz.restype = ctypes.c_longlong
        self.L.im_ib_avg_len.restype = ctypes.c_double
        self.L.im_ib_term_bytes.restype = ctypes.c_longlong
        self.h = ctypes.c_void_p(self.L.im_ib_create())

    def close(self):
        if self.h:
            self.L.im_ib_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass

    def add_text(self, text: str) -> int:
        raw = text.encode("utf-8")
        return self.L.im_ib_add_text(self.h, raw, ctypes.c_longlong(len(raw)))

    def add_terms(self, ids) -> int:
        arr = np.ascontiguousarray(ids, dtype=np.int32)
        return self.L.im_ib_add_terms(self.h, arr.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(arr.size))

    def tokenize(self, text: str, add: bool = False) -> np.ndarray:
        raw = text.encode("utf-8")
        cap = max(8, len(raw) // 1 + 1)
        out = np.empty(cap, dtype=np.int32)
        n = self.L.im_ib_tokenize(self.h, raw, ctypes.c_longlong(len(raw)), ctypes.c_int(1 if add else 0),
                                  out.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(cap))
        return out[:min(n, cap)].copy()

    def lookup(self, term: str) -> int:
        raw = term.encode("utf-8")
        return self.L.im_ib_lookup(self.h, raw, ctypes.c_longlong(len(raw)))

    def term(self, tid: int) -> str:
        buf = ctypes.create_string_buffer(256)
        n = self.L.im_ib_term_bytes(self.h, ctypes.c_int(tid), buf, ctypes.c_longlong(256))
        return buf.raw[:max(0, min(n, 256))].decode("utf-8", "replace")

    @property
    def vocab(self) -> int:
        return int(self.L.im_ib_vocab(self.h))

    @property
    def n_docs(self) -> int:
        return int(self.L.im_ib_docs(self.h))

    def export(self):
        """-> dict(off int64[V+1], doc int32[nnz], tf uint8[nnz], doc_len int32[n], df int32[V])"""
        V, n, nnz = self.vocab, self.n_docs, int(self.L.im_ib_nnz(self.h))
        off = np.empty(V + 1, np.int64)
        doc = np.empty(max(nnz, 1), np.int32)
        tf = np.empty(max(nnz, 1), np.uint8)
        dl = np.empty(max(n, 1), np.int32)
        df = np.empty(max(V, 1), np.int32)
        self.L.im_ib_export(self.h, off.ctypes.data_as(ctypes.c_void_p), doc.ctypes.data_as(ctypes.c_void_p),
                            tf.ctypes.data_as(ctypes.c_void_p), dl.ctypes.data_as(ctypes.c_void_p),
                            df.ctypes.data_as(ctypes.c_void_p))
        return dict(off=off, doc=doc[:nnz], tf=tf[:nnz], doc_len=dl[:n], df=df[:V])


def bm25_ref(csr: dict, q_terms, n_docs_global=None, avg_len=None, df_global=None, k=10, alive=None):
    """NumPy oracle: AND semantics, FTS5 BM25; returns list of (score, doc) sorted (score desc, doc asc)."""
    off, doc, tf, dl = csr["off"], csr["doc"], csr["tf"], csr["doc_len"]
    n = len(dl)
    N = n if n_docs_global is None else n_docs_global
    avg = (dl.sum() / max(n, 1)) if avg_len is None else avg_len
    df = csr["df"] if df_global is None else df_global
    terms = []
    for t in q_terms:
        if t == -1:
            continue
        if t < 0 or t >= len(off) - 1:
            return []
        if t not in terms:
            terms.append(int(t))
    if not terms:
        return []
    idf = idf_fts5(N, np.asarray([df[t] for t in terms]))
    cand = None
    for t in terms:
        d = doc[off[t]:off[t + 1]]
        cand = d if cand is None else np.intersect1d(cand, d, assume_unique=True)
    if cand is None or len(cand) == 0:
        return []
    scores = np.zeros(len(cand), dtype=np.float64)
    norm = K1 * (1 - B + B * dl[cand].astype(np.float64) / max(avg, 1e-9))
    for w, t in zip(idf, terms):
        d = doc[off[t]:off[t + 1]]
        f = tf[off[t]:off[t + 1]]
        pos = np.searchsorted(d, cand)
        tfv = f[pos].astype(np.float64)
        scores += float(w) * tfv * (K1 + 1) / (tfv + norm)
    if alive is not None:
        keep = alive[cand] != 0
        cand, scores = cand[keep], scores[keep]
    order = np.lexsort((cand, -scores))[:k]
    return [(float(scores[i]), int(cand[i])) for i in order]


class Bm25Index:
    """One shard's CSR postings resident in HBM + the query kernel."""

    def __init__(self, csr: dict, device="cuda", n_docs_global=None, avg_len=None, df_global=None):
        import torch

        self.device = torch.device(device)
        dl = csr["doc_len"]
        self.n_docs = len(dl)
        self.vocab = len(csr["off"]) - 1
        N = self.n_docs if n_docs_global is None else n_docs_global
        self.avg_len = float(dl.sum() / max(self.n_docs, 1)) if avg_len is None else float(avg_len)
        df = csr["df"] if df_global is None else df_global
        t = lambda a, dt: torch.as_tensor(np.ascontiguousarray(a), dtype=dt).to(self.device)  # noqa: E731
        self.off = t(csr["off"], torch.int64)
        self.doc = t(csr["doc"] if len(csr["doc"]) else np.zeros(1, np.int32), torch.int32)
        self.tf = t(csr["tf"] if len(csr["tf"]) else np.zeros(1, np.uint8), torch.uint8)
        self.norm = t(K1 * (1 - B + B * dl.astype(np.float32) / max(self.avg_len, 1e-9)), torch.float32)
        self.idf = t(idf_fts5(N, df), torch.float32)
        self.dense_slot = self.dense_tf = None
        self.build_dense_maps()

    def build_dense_maps(self, min_df: int | None = None, max_bytes: int = 8 << 30) -> int:
        """Direct tf maps for terms whose posting list covers >= 1/32 of the shard (at least 4096 documents): probing them
        in an intersection becomes one byte load.  Returns the number of dense terms (0 = none, kernel takes the plain path)."""
        import torch

        n, V = self.n_docs, self.vocab
        if n == 0 or V == 0:
            return 0
        min_df = max(4096, n // 32) if min_df is None else min_df
        df_local = (self.off[1:] - self.off[:-1])
        dense = torch.nonzero(df_local >= min_df).flatten()
        if dense.numel() == 0 or dense.numel() * n > max_bytes:
            dense = dense[torch.argsort(df_local[dense], descending=True)][:max(0, max_bytes // max(n, 1))]
        if dense.numel() == 0:
            self.dense_slot = self.dense_tf = None
            return 0
        slot = torch.full((V,), -1, dtype=torch.int32, device=self.device)
        slot[dense] = torch.arange(dense.numel(), dtype=torch.int32, device=self.device)
        tfmap = torch.zeros((dense.numel(), n), dtype=torch.uint8, device=self.device)
        for s_i, t_id in enumerate(dense.tolist()):
            a, b = int(self.off[t_id]), int(self.off[t_id + 1])
            tfmap[s_i, self.doc[a:b].long()] = self.tf[a:b].clamp(min=1)
        self.dense_slot, self.dense_tf = slot, tfmap
        return int(dense.numel())

    @classmethod
    def from_device_csr(cls, off, doc, tf, norm, idf, avg_len):
        """Adopt CSR tensors already built on the GPU (synthetic corpus generator)."""
        self = cls.__new__(cls)
        self.device = off.device
        self.off, self.doc, self.tf, self.norm, self.idf = off, doc, tf, norm, idf
        self.n_docs = norm.numel()
        self.vocab = off.numel() - 1
        self.avg_len = float(avg_len)
        self.dense_slot = self.dense_tf = None
        self.build_dense_maps()
        return self

    def nbytes(self) -> int:
        base = sum(x.numel() * x.element_size() for x in (self.off, self.doc, self.tf, self.norm, self.idf))
        return base + (self.dense_tf.numel() + self.dense_slot.numel() * 4 if self.dense_tf is not None else 0)

    def search_partials(self, q_terms, alive=None, blocks_per_query=0, counts=None):
        """q_terms: int32 [nq, T<=16] (-1 padded) on device -> per-warp lists (scores[P,nq,32], ids[P,nq,32])."""
        import torch

        nq, T = q_terms.shape
        assert q_terms.dtype == torch.int32 and T <= MAX_TERMS and q_terms.is_contiguous()
        L = _native.require()
        if blocks_per_query <= 0:
            # 16 CTAs (64 warps) per query: a query whose driving list has millions of postings is spread over enough
            # warps, and the warps of a selective query find their slice empty and exit at once
            blocks_per_query = max(1, min(64, max(16, (2 * L.im_sm_count()) // max(nq, 1))))
        P = blocks_per_query * 4
        out_s = torch.empty((P, nq, 32), device=self.device, dtype=torch.float32)
        out_i = torch.empty((P, nq, 32), device=self.device, dtype=torch.int32)
        if self.dense_tf is not None:
            rc = L.im_bm25_topk_dense(_native.ptr(self.off), _native.ptr(self.doc), _native.ptr(self.tf), _native.ptr(self.norm),
                                      _native.ptr(self.idf), _native.ptr(alive), _native.ptr(q_terms), ctypes.c_int(nq),
                                      ctypes.c_int(T), ctypes.c_int(self.vocab), ctypes.c_int(blocks_per_query),
                                      _native.ptr(out_s), _native.ptr(out_i), _native.ptr(counts), _native.ptr(self.dense_slot),
                                      _native.ptr(self.dense_tf), ctypes.c_longlong(self.n_docs), _native.stream_ptr())
        else:
            rc = L.im_bm25_topk(_native.ptr(self.off), _native.ptr(self.doc), _native.ptr(self.tf), _native.ptr(self.norm),
                                _native.ptr(self.idf), _native.ptr(alive), _native.ptr(q_terms), ctypes.c_int(nq),
                                ctypes.c_int(T), ctypes.c_int(self.vocab), ctypes.c_int(blocks_per_query),
                                _native.ptr(out_s), _native.ptr(out_i), _native.ptr(counts), _native.stream_ptr())
        if rc < 0:
            _native.check(rc, "im_bm25_topk")
        _native.count_launch()
        return out_s, out_i

    def search(self, q_terms, k=20, alive=None, id_offset=0, push=None):
        """``push``: a ``TopkChannel`` -- the merge kernel also stores the shard's list into every peer's receive area."""
        from infomesh_b200.ops.search import topk_merge

        assert k <= 32
        ps, pi = self.search_partials(q_terms, alive)
        return topk_merge(ps, pi, k, id_offset=id_offset, push=push)


def passage_score_ref(tokens, pass_bounds, q_terms):
    """Oracle for one (query, doc): returns (best_score, best_passage) with coverage + 0.1 * density."""
    q = [t for t in dict.fromkeys(int(x) for x in q_terms) if t >= 0]
    if not q:
        return 0.0, -1
    best, bp = -1.0, -1
    for p in range(len(pass_bounds) - 1):
        seg = tokens[pass_bounds[p]:pass_bounds[p + 1]]
        if len(seg) == 0:
            s = 0.0
        else:
            present = sum(1 for t in q if t in set(int(x) for x in seg))
            hits = sum(1 for x in seg if int(x) in q)
            s = np.float32(present) / np.float32(len(q)) + np.float32(0.1) * np.float32(hits) / np.float32(len(seg))
        if s > best:
            best, bp = float(s), p
    return (best if bp >= 0 else 0.0), bp


def passage_score(tok, pass_off, doc_pass_off, pair_doc, pair_query, q_terms):
    """CUDA passage selection for (query, doc) pairs; see csrc/search/bm25.cu ``passage_score_kernel``."""
    import torch

    n_pairs = pair_doc.numel()
    nq, T = q_terms.shape
    out_s = torch.empty((n_pairs,), device=tok.device, dtype=torch.float32)
    out_p = torch.empty((n_pairs,), device=tok.device, dtype=torch.int32)
    L = _native.require()
    rc = L.im_passage_score(_native.ptr(tok), _native.ptr(pass_off), _native.ptr(doc_pass_off), _native.ptr(pair_doc),
                            _native.ptr(pair_query), _native.ptr(q_terms), ctypes.c_int(n_pairs), ctypes.c_int(T),
                            _native.ptr(out_s), _native.ptr(out_p), _native.stream_ptr())
    _native.check(rc, "im_passage_score")
    _native.count_launch()
    return out_s, out_p


_ = math
