"""End-to-end timing of the PRODUCT search surfaces: text in -> formatted answer out, including query
tokenisation, the device pass, SQLite document fetch, snippet cutting and response formatting.

    python scripts/bench_product_search.py [--docs 50000] [--queries 200] [--gpus 1]

Paths timed on the same store:
  * MCP ``web_search`` (``mcp/handlers.ToolRuntime.web_search``, the call the MCP server and the HTTP /search route make),
    one query at a time (interactive latency), query cache off;
  * SDK-level ``GpuSearchIndex.search_many`` with 64-query batches (throughput);
  * the CPU path (``search_local`` over SQLite FTS5) for the same queries, for scale.
Weights are random-init (no checkpoints offline), so ``allow_untrained_models`` is set: this measures time, not quality."""
import argparse
import asyncio
import json
import statistics
import sys
import tempfile
import time
from dataclasses import replace
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=50000)
    ap.add_argument("--queries", type=int, default=200)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    from infomesh_b200.config import Config
    from infomesh_b200.mcp.handlers import ToolRuntime
    from infomesh_b200.search import query as Q
    from infomesh_b200.services import AppContext

    tmp = Path(tempfile.mkdtemp(prefix="im_prod_"))
    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp), index=replace(base.index, db_path=tmp / "index.db", vector_search=False),
                  gpu=replace(base.gpu, enabled=True, devices=a.gpus, allow_untrained_models=True, query_batch=64))
    ctx = AppContext(cfg)
    rng = np.random.default_rng(5)
    vocab = [f"w{i}" for i in range(20000)]
    p = 1.0 / np.arange(1, len(vocab) + 1) ** 1.05
    p /= p.sum()
    t0 = time.time()
    ctx.store._conn.execute("BEGIN")
    for i in range(a.docs):
        words = rng.choice(len(vocab), size=int(rng.integers(60, 200)), p=p)
        text = " ".join(vocab[w] for w in words)
        ctx.store._conn.execute("INSERT INTO documents (url, title, text, raw_html_hash, text_hash, crawled_at) VALUES (?,?,?,?,?,?)",
                                (f"https://site{i % 211}.example/p/{i}", " ".join(vocab[w] for w in words[:5]), text, f"r{i}", f"t{i}", time.time() - float(rng.integers(0, 86400 * 30))))
    ctx.store._conn.execute("COMMIT")
    t_store = time.time() - t0
    from infomesh_b200.mcp.server import attach_gpu_index

    t0 = time.time()
    gi = attach_gpu_index(ctx)
    t_build = time.time() - t0
    assert gi is not None, "GPU index did not come up"
    rt = ToolRuntime(ctx)
    queries = [" ".join(vocab[w] for w in rng.choice(len(vocab), size=int(rng.integers(1, 4)), p=p)) for _ in range(a.queries)]

    async def mcp_loop():
        lat = []
        for q in queries[:8]:
            await rt.web_search({"query": q, "top_k": 5})
        for q in queries:
            rt.query_cache.clear()
            t = time.perf_counter()
            out = await rt.web_search({"query": q, "top_k": 5})
            lat.append((time.perf_counter() - t) * 1e3)
        return lat, out

    lat, sample = asyncio.run(mcp_loop())
    gi.search_many(queries[:64], 10)
    t = time.perf_counter()
    n_done = 0
    for rep in range(3):
        res = gi.search_many(queries, 10)
        n_done += len(res)
    t_batch = time.perf_counter() - t
    t = time.perf_counter()
    for q in queries[:50]:
        Q.search_local(ctx.store, q, limit=10)
    t_cpu = (time.perf_counter() - t) / 50 * 1e3
    out = {"docs": a.docs, "gpus": a.gpus, "store_build_s": round(t_store, 1), "gpu_index_build_s": round(t_build, 1), "index": gi.stats() if a.gpus == 1 else {k: v for k, v in gi.stats().items() if k != "per_rank"},
           "mcp_web_search_ms": {"p50": round(statistics.median(lat), 3), "p95": round(sorted(lat)[int(0.95 * len(lat))], 3), "mean": round(sum(lat) / len(lat), 3), "n": len(lat)},
           "search_many_batch64_qps": round(n_done / t_batch, 1), "cpu_search_local_ms_per_query": round(t_cpu, 3),
           "sample_answer_chars": len(sample), "note": "random-init models (allow_untrained_models): time only"}
    out["index"].pop("health", None)
    print(json.dumps(out))
    closer = getattr(gi, "close", None)
    if closer:
        closer()
    ctx.close()


if __name__ == "__main__":
    main()
