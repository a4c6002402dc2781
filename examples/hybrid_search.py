"""The fused GPU pipeline over a LocalStore: encoder -> dense top-k + BM25 -> RRF -> cross-encoder rerank -> top-10,
one CUDA-graph replay per batch of queries.  Needs a B200 (sm_100a); weights are random-initialised offline."""
import tempfile
import time
from pathlib import Path

import torch

from infomesh_b200.engine.gpu_index import GpuSearchIndex
from infomesh_b200.index.local_store import LocalStore

TOPICS = ["tensor memory accumulators on blackwell", "kademlia routing table buckets", "sqlite full text search ranking",
          "merkle tree audit proofs", "simhash near duplicate detection", "token bucket bandwidth throttling"]

assert torch.cuda.is_available(), "this example needs a CUDA device"
with tempfile.TemporaryDirectory() as d:
    store = LocalStore(Path(d) / "index.db")
    for i in range(600):
        t = TOPICS[i % len(TOPICS)]
        store.add_document(url=f"https://example.org/{i}", title=f"{t.title()} #{i}", text=f"{t} — note {i}. " * 8, raw_html_hash=f"r{i}", text_hash=f"t{i}", language="en")
    # allow_untrained: this demo ships no checkpoints, so the (random-init) encoder / reranker are allowed to rank; a real
    # deployment passes encoder_path= / reranker_path= instead and leaves the flag off
    index = GpuSearchIndex(store, query_batch=16, allow_untrained=True)
    print("resident documents:", index.rebuild(), index.stats())
    queries = ["kademlia buckets", "merkle audit proofs", "bandwidth throttling", "duplicate detection"]
    index.search_many(queries)                       # warm-up (captures the CUDA graph)
    t0 = time.perf_counter()
    results = index.search_many(queries, k=3)
    print(f"{len(queries)} queries in {(time.perf_counter() - t0) * 1e3:.2f} ms")
    for q, hits in zip(queries, results):
        print(q, "->", [h["title"] for h in hits])
    store.close()
