#!/usr/bin/env python
"""Reference arm of bench.py: the UNMODIFIED reference (baseline/_ref/infomesh) answering the same query stream
through its own public API and stock code path.

    LocalStore.add_document(...)                    infomesh/index/local_store.py:198-251   (index build, untimed)
    VectorStore(...)._collection.upsert(...)        infomesh/index/vector_store.py:75-95    (index build, untimed)
    search_hybrid(store, vector_store, q, limit=10) infomesh/search/query.py:244-319        (TIMED, one call per query)
        -> LocalStore.search  (SQLite FTS5 MATCH + bm25())      local_store.py:253-352
        -> VectorStore.search (SentenceTransformer.encode + collection.query)  vector_store.py:187-254
        -> merge_results      (RRF k=60)                        search/merge.py:37-133

Nothing of infomesh_b200 (models, kernels, engine) is imported here.  The corpus is the same counter-hash Zipfian
corpus as the repo arm (same generator constants; `tests/test_reference_arm.py` checks the two generators agree),
rendered as text: term id t -> word "w<t>".  Document vectors are the same random unit vectors (bulk-upserted through
the ChromaDB collection API instead of 1-by-1 `VectorStore.add_document`, which would encode every document with a
batch-1 forward pass); QUERIES are encoded by the model inside the timed region exactly as the reference does.

Departures from the repo arm's config, all reported in the JSON line (`same_config: false`):
  * corpus size: `--ref-docs` (default 1M) instead of 10M -- the reference indexes through per-row SQLite INSERT+COMMIT
    and a Python dict per vector; 10M does not fit the driver's time box.  Fewer documents make every reference query
    CHEAPER (shorter postings, smaller similarity scan), so the ratio computed from this arm favours the reference.
  * no reranker: the reference's reranker is an LLM prompt sent to an external ollama/llama.cpp/vLLM HTTP server
    (infomesh/search/reranker.py:86-163); none exists offline, and `rerank_with_llm` keeps the input order on failure.
    The repo arm's headline DOES include its cross-encoder; its `retrieval_only` key is the like-for-like number.
  * third-party wheels missing from the image are replaced by library-only stand-ins in baseline/shims (structlog ->
    stdlib logging; chromadb/hnswlib -> exact torch matmul+topk on the GPU; sentence-transformers -> HF transformers
    BertModel, random-init, fp32, CUDA when visible)."""
from __future__ import annotations

import json
import os
import statistics
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent

# ---- the repo arm's corpus generator, restated with plain torch-on-CPU (engine/synth.py constants) ----
VOCAB_TERMS, DOC_LEN, DIM, ZIPF_S, SEED = 200_000, 64, 384, 1.0, 1234
_M1 = 6364136223846793005
_M2 = -4417276706812531889
_INC = 1442695040888963407


def _hash_uniform(idx, seed):
    import torch

    x = idx * _M1 + ((seed * _INC) & 0x7FFFFFFFFFFFFFFF)
    x = x ^ ((x >> 33) & 0x7FFFFFFF)
    x = x * _M2
    x = x ^ ((x >> 29) & 0x7FFFFFFFF)
    x = x * _M1
    x = x ^ ((x >> 32) & 0xFFFFFFFF)
    return ((x >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / 16777216.0)


def _zipf_cdf():
    import torch

    w = 1.0 / torch.arange(1, VOCAB_TERMS + 1, dtype=torch.float64).pow(ZIPF_S)
    return (w.cumsum(0) / w.sum()).to(torch.float32)


def doc_terms(start: int, count: int, cdf):
    import torch

    rows = torch.arange(start, start + count, dtype=torch.int64)
    col = torch.arange(DOC_LEN, dtype=torch.int64)[None, :]
    u = _hash_uniform(rows[:, None] * DOC_LEN + col, SEED)
    return torch.searchsorted(cdf, u).clamp_(max=VOCAB_TERMS - 1)


def doc_vectors(start: int, count: int, device="cpu"):
    import torch

    rows = torch.arange(start, start + count, dtype=torch.int64, device=device)
    col = torch.arange(DIM, dtype=torch.int64, device=device)[None, :]
    idx = rows[:, None] * DIM + col
    u1 = _hash_uniform(idx, SEED + 17).clamp_(min=1e-7)
    u2 = _hash_uniform(idx, SEED + 31)
    v = torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(6.283185307179586 * u2)
    return torch.nn.functional.normalize(v, dim=1)


def query_terms(n_docs_global: int, n_queries: int, cdf, seed: int = 99, n_terms=(2, 3), mix: str = "rare"):
    """Same construction as engine/synth.make_queries: 2-3 terms of a random document.  ``mix="rare"`` spreads the picks
    over the 45%..90% rarity quantiles; ``"common"`` additionally forces one of the document's most frequent terms."""
    import torch

    g = torch.Generator(device="cpu").manual_seed(seed)
    doc_ids = torch.randint(0, n_docs_global, (n_queries,), generator=g)
    k_terms = torch.randint(n_terms[0], n_terms[1] + 1, (n_queries,), generator=g)
    col = torch.arange(DOC_LEN, dtype=torch.int64)
    out = []
    for i in range(n_queries):
        d = int(doc_ids[i])
        u = _hash_uniform(d * DOC_LEN + col, SEED)
        row = torch.searchsorted(cdf, u).clamp_(max=VOCAB_TERMS - 1)
        terms = torch.unique(row).sort().values
        kt = min(int(k_terms[i]), terms.numel())
        pos = [int(round((terms.numel() - 1) * (0.45 + 0.45 * j / max(kt - 1, 1)))) for j in range(kt)]
        if mix == "common" and i % 2 == 1:
            pos[0] = 0
        out.append([int(terms[p]) for p in sorted(set(pos))])
    return out, doc_ids


def _word(t: int) -> str:
    return f"w{t}"


def run(args, ClockSampler=None) -> int:
    if int(os.environ.get("RANK", "0")) != 0:
        return 0                                   # the reference is a single-process program: rank 0 runs it
    ref_dir = HERE / "_ref"
    if not (ref_dir / "infomesh").is_dir():
        try:
            sys.path.insert(0, str(HERE))
            import install_ref

            install_ref.install()
        except Exception:  # noqa: BLE001
            pass
    if not (ref_dir / "infomesh").is_dir():
        print(json.dumps({"impl": "reference", "unavailable": "baseline/_ref/infomesh missing and /root/reference not "
                          "mounted (run python baseline/install_ref.py where the reference is available)"}))
        return 0
    sys.path.insert(0, str(ref_dir))
    sys.path.append(str(HERE / "shims"))           # AFTER site-packages: real wheels win when they exist
    import logging

    logging.getLogger("infomesh").setLevel(logging.ERROR)
    try:
        import torch
        from infomesh.index.local_store import LocalStore
        from infomesh.index.vector_store import VectorStore
        from infomesh.search.query import search_hybrid
    except Exception as exc:  # noqa: BLE001
        print(json.dumps({"impl": "reference", "unavailable": f"reference import failed: {type(exc).__name__}: {exc}"}))
        return 0

    n_docs = int(args.ref_docs)
    B, K, W = args.batch, args.steps, args.warmup
    cdf = _zipf_cdf()
    t0 = time.time()
    store = LocalStore()                                       # in-memory SQLite + FTS5, stock defaults
    vs = VectorStore(model_name="BAAI/bge-small-en")           # EphemeralClient, cosine space
    chunk = 20_000
    budget = float(getattr(args, "ref_build_budget_s", 150.0))
    built = 0
    for a in range(0, n_docs, chunk):
        if time.time() - t0 > budget:
            break
        b = min(n_docs, a + chunk)
        terms = doc_terms(a, b - a, cdf).tolist()
        ids, metas = [], []
        for j, row in enumerate(terms):
            gid = a + j
            url = f"https://d{gid}.example/p"
            title = f"doc {gid}"
            text = " ".join(map(_word, row))
            did = store.add_document(url, title, text, f"r{gid:x}", f"t{gid:x}")
            ids.append(str(did))
            metas.append({"url": url, "title": title, "text_preview": text[:500]})
        vs._collection.upsert(ids=ids, embeddings=doc_vectors(a, b - a), metadatas=metas)
        built = b
    n_docs = built                                              # time-boxed build: report what was actually indexed
    build_s = time.time() - t0
    qterms, _ = query_terms(n_docs, (K + W) * B, cdf, mix=args.query_mix)
    queries = [" ".join(_word(t) for t in q) for q in qterms]
    cuda = torch.cuda.is_available()

    def step(i):
        out = None
        for q in queries[i * B:(i + 1) * B]:
            out = search_hybrid(store, vs, q, limit=10)
        return out

    for i in range(W):
        last = step(i)
    if cuda:
        torch.cuda.synchronize()
    sampler = None
    if ClockSampler is not None and cuda:
        sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
        sampler.start()
    per = []
    t_all = time.perf_counter()
    for i in range(K):
        t1 = time.perf_counter()
        last = step(W + i)
        per.append((time.perf_counter() - t1) * 1e3)
    if cuda:
        torch.cuda.synchronize()
    total_s = time.perf_counter() - t_all
    clocks = sampler.stop() if sampler is not None else {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no GPU sampled"]}
    qps = B * K / total_s
    n_hybrid = sum(1 for m in last.results if m.fts_score is not None) if last is not None else 0
    seq = 8                                                     # [CLS] + 2-3 words + [SEP], padded per call
    line = {
        "impl": "reference",
        "metric": "queries/sec, hybrid BM25+dense top-10 over a synthetic index (reference search_hybrid, no reranker)",
        "value": round(qps, 2), "unit": "queries/s", "n_gpus": int(getattr(args, "gpus", 1) or 1), "gpus_used": 1, "steps": K, "warmup": W,
        "gpus_note": "the reference has no multi-GPU path: launched with N ranks, rank 0 runs its stock single-process search and the "
                     "other ranks idle, so `value` is the same at every N",
        "ms_per_step": round(total_s * 1e3 / K, 3), "p50_step_ms": round(statistics.median(per), 3),
        "p50_query_ms": round(statistics.median(per) / B, 4),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "fp32 (sentence-transformers default) on " + ("cuda" if cuda else "cpu"),
        "data": "synthetic (same Zipfian corpus generator and random unit vectors as the repo arm, rendered as text; "
                "random-init weights)",
        "same_config": False,
        "config": {
            "model": "bge-small-en architecture via HF transformers (random-init), no reranker (reference reranker = "
                     "LLM over HTTP, unavailable offline)",
            "index_docs": n_docs, "repo_arm_index_docs": 10_000_000, "dim": DIM, "global_batch": B, "top_k": 10,
            "query_mix": args.query_mix,
            "parallelism": "single process, one query per search_hybrid() call (the reference has no batch or multi-GPU path)",
            "api": "infomesh.search.query.search_hybrid(LocalStore, VectorStore, q, limit=10)",
            "stand_ins": "structlog->logging, chromadb->exact torch topk (GPU), sentence-transformers->transformers.BertModel",
            "index_build_s": round(build_s, 1),
            "l2_policy": "a distinct query batch every step; SQLite pages and the vector matrix exceed L2",
            "last_query_fts_hits_in_top10": n_hybrid,
        },
        "e2e": {"value": round(qps, 2), "unit": "queries/s", "ms_per_step": round(total_s * 1e3 / K, 3),
                "note": "the reference API is text-in / objects-out: this IS its end-to-end number",
                "h2d_bytes_per_step": B * (2 * seq * 8 + DIM * 4) if cuda else 0,
                "d2h_bytes_per_step": B * (DIM * 4 + 10 * 12) if cuda else 0},
        "gpu_launches": 0,
        "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    return 0
