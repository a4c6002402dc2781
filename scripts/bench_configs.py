"""The other BASELINE.json configurations, measured on ONE GPU (bench.py is the batch-64 hybrid headline):

  1. LocalStore BM25 search over 1k synthetic docs on CPU                       -> queries/s
  2. dense vector search, bge-small (random init), 10M x 384 index, batch-1     -> p50 latency (CUDA graph replay)
  4. RAG path: best-passage extraction + t5-small summariser (random init)      -> generated tokens/s
  5. index build: encode (bge-small) + SimHash fingerprints + near-dup scan     -> passages/s

Prints one JSON object.  `--cpu-only` runs only (1).  Everything is synthetic / random-init (no network here)."""
import json
import random
import statistics
import sys
import time

out = {}

# ------------------------------------------------------------------ 1. CPU BM25
from infomesh_b200.hashing import content_hash
from infomesh_b200.index.local_store import LocalStore

rng = random.Random(0)
vocab = [f"w{i}" for i in range(4000)]
store = LocalStore(None)
for d in range(1000):
    text = " ".join(rng.choices(vocab, k=200))
    store.add_document(f"https://ex.org/{d}", f"doc {d}", text, content_hash(f"h{d}"), content_hash(text))
qs = [" ".join(rng.choices(vocab, k=2)) for _ in range(500)]
for q in qs[:20]:
    store.search(q, limit=10)
t0 = time.perf_counter()
for q in qs:
    store.search(q, limit=10)
dt = time.perf_counter() - t0
out["localstore_bm25_1k_cpu"] = {"queries_per_s": round(len(qs) / dt, 1), "p_mean_ms": round(dt / len(qs) * 1e3, 3)}

if "--cpu-only" in sys.argv:
    print(json.dumps(out))
    sys.exit(0)

import torch

from infomesh_b200.models.bert import BGE_SMALL, BertModel
from infomesh_b200.models.t5 import T5_SMALL, T5Model
from infomesh_b200.ops import dedup as DD
from infomesh_b200.ops.search import sim_topk

dev = torch.device("cuda:0")


def graph_p50(fn, n=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


# ------------------------------------------------------------------ 2. dense batch-1 top-10 over 10M
try:
    n_docs = 10_000_000
    enc = BertModel(BGE_SMALL, device=dev, seed=1)
    docs = torch.empty((n_docs, 384), device=dev, dtype=torch.bfloat16)
    for a in range(0, n_docs, 1_000_000):
        docs[a:a + 1_000_000] = torch.nn.functional.normalize(torch.randn(1_000_000, 384, device=dev), dim=1).bfloat16()
    ids = torch.randint(1000, 20000, (1, 32), dtype=torch.int32, device=dev)
    lens = torch.tensor([9], dtype=torch.int32, device=dev)
    ms_all = graph_p50(lambda: sim_topk(enc.embed(ids, lens), docs, 10))
    q = enc.embed(ids, lens)
    ms_search = graph_p50(lambda: sim_topk(q, docs, 10))
    out["dense_batch1_top10_10M"] = {"p50_ms": round(ms_all, 3), "search_only_ms": round(ms_search, 3),
                                     "search_GBps": round(n_docs * 768 / ms_search / 1e6, 1)}
    del docs
except Exception as exc:  # noqa: BLE001
    out["dense_batch1_top10_10M"] = {"error": repr(exc)}

# ------------------------------------------------------------------ 4. RAG: passage extraction + t5-small
try:
    from infomesh_b200.search.rag import extract_answers  # noqa: F401  (CPU passage extraction lives in search/)
    t5 = T5Model(T5_SMALL, device=dev, seed=3)
    B, S, new = 16, 256, 32
    ids = torch.randint(5, 30000, (B, S), dtype=torch.int32, device=dev)
    lens = torch.full((B,), S, dtype=torch.int32, device=dev)
    for _ in range(2):
        t5.generate(ids, lens, max_new_tokens=new, check_every=new)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    reps = 3
    for _ in range(reps):
        o = t5.generate(ids, lens, max_new_tokens=new, check_every=new)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    out["rag_t5_small_summarise"] = {"batch": B, "input_tokens": S, "new_tokens": new, "ms_per_batch": round(ms, 2),
                                     "generated_tokens_per_s": round(B * o.shape[1] / ms * 1e3, 1)}
except Exception as exc:  # noqa: BLE001
    out["rag_t5_small_summarise"] = {"error": repr(exc)}

# ------------------------------------------------------------------ 5. index build: encode + SimHash + near-dup scan
try:
    enc = BertModel(BGE_SMALL, device=dev, seed=1)
    Bp, Sp = 4096, 128
    ids = torch.randint(1000, 20000, (Bp, Sp), dtype=torch.int32, device=dev)
    lens = torch.randint(64, Sp + 1, (Bp,), dtype=torch.int32, device=dev)
    ms_enc = graph_p50(lambda: enc.embed(ids, lens), n=10)
    texts = [" ".join(rng.choices(vocab, k=120)) for _ in range(Bp)]
    text, ws, we, off = DD.normalize_batch(texts)
    tt, tws, twe, toff = (torch.from_numpy(x).to(dev) for x in (text, ws, we, off))
    table = torch.randint(-2**62, 2**62, (10_000_000,), dtype=torch.int64, device=dev)   # fingerprints already indexed

    def fp_and_scan():
        fp = DD.simhash_from_arrays(tt, tws, twe, toff)
        return DD.hamming_scan(table, fp)

    ms_dd = graph_p50(fp_and_scan, n=10)
    out["index_build_encode_simhash"] = {"batch_passages": Bp, "tokens_per_passage": Sp, "encode_ms": round(ms_enc, 3),
                                         "simhash_plus_scan10M_ms": round(ms_dd, 3),
                                         "passages_per_s": round(Bp / (ms_enc + ms_dd) * 1e3, 1),
                                         "encode_only_passages_per_s": round(Bp / ms_enc * 1e3, 1)}
except Exception as exc:  # noqa: BLE001
    out["index_build_encode_simhash"] = {"error": repr(exc)}

print(json.dumps(out))
