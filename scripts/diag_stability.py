#!/usr/bin/env python3
"""Long-running stability monitor: drive a node's hot paths in a loop and watch RSS, open descriptors, thread count,
Python heap and (when present) device memory for drift (counterpart of reference scripts/diag_stability.py, which
watches the dashboard; here the watched workload is selectable).

    python scripts/diag_stability.py --duration 300 --interval 5                  # local search loop on a synthetic index
    python scripts/diag_stability.py --workload dashboard --duration 3600         # dashboard data providers
    python scripts/diag_stability.py --workload gpu --duration 600                # hybrid engine (needs a B200)

Exit code 1 when any resource grew by more than ``--tolerance`` (default 10 %) between the first and last quarter."""
from __future__ import annotations

import argparse
import gc
import os
import statistics
import sys
import tempfile
import threading
import time
import tracemalloc
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


def _sample(gpu: bool) -> dict[str, float]:
    import psutil

    p = psutil.Process()
    out = {"rss_mb": p.memory_info().rss / 2 ** 20, "fds": float(p.num_fds()), "threads": float(threading.active_count()),
           "py_heap_mb": tracemalloc.get_traced_memory()[0] / 2 ** 20}
    if gpu:
        import torch

        out["cuda_alloc_mb"] = torch.cuda.memory_allocated() / 2 ** 20
        out["cuda_reserved_mb"] = torch.cuda.memory_reserved() / 2 ** 20
    return out


def _search_workload(tmp: Path):
    from infomesh_b200.crawler.parser import ParsedPage
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.search.query import search_local
    from infomesh_b200.services import index_document

    store = LocalStore(tmp / "index.db")
    topics = ["tensor memory", "thread block clusters", "bulk tensor copies", "peer memory collectives", "cuda graphs"]
    for i in range(400):
        t = topics[i % len(topics)]
        index_document(ParsedPage(url=f"https://diag.example/{i}", title=f"{t} {i}", text=f"Document {i} discusses {t} in depth. " * 12,
                                  language="en", raw_html_hash=f"r{i}", text_hash=f"t{i}"), store)
    it = iter(range(10 ** 12))
    return lambda: search_local(store, topics[next(it) % len(topics)], limit=10)


def _dashboard_workload(tmp: Path):
    from dataclasses import replace

    from infomesh_b200.config import Config
    from infomesh_b200.dashboard.text_report import render_text_report

    base = Config()
    cfg = replace(base, node=replace(base.node, data_dir=tmp), index=replace(base.index, db_path=tmp / "index.db", vector_search=False))
    return lambda: render_text_report(cfg)


def _gpu_workload(_tmp: Path):
    import torch

    from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
    from infomesh_b200.engine.synth import SynthConfig, SynthShard, make_queries

    dev = torch.device("cuda:0")
    n = 500_000
    shard = SynthShard(SynthConfig(n_docs=n, n_docs_global=n), device=dev)
    cfg = HybridConfig(nq=32)
    eng = HybridEngine(shard, cfg, docs_per_shard=n)
    q_terms, q_tok, q_len, _ = make_queries(SynthConfig(n_docs=n, n_docs_global=n), 32, max_terms=cfg.max_terms,
                                            max_q_tokens=cfg.max_q_tokens, device=dev)
    enc_ids = torch.zeros((32, cfg.enc_seq), dtype=torch.int32, device=dev)
    L = min(cfg.enc_seq - 2, q_tok.shape[1])
    enc_ids[:, 0], enc_ids[:, 1:1 + L] = 101, (1000 + q_tok[:, :L].long() % 20000).to(torch.int32)
    enc_len = (q_len.clamp(max=L) + 2).to(torch.int32)
    host_s = torch.empty((32, cfg.k_out), dtype=torch.float32).pin_memory()
    host_i = torch.empty((32, cfg.k_out), dtype=torch.int64).pin_memory()
    batch = tuple(x.cpu().pin_memory() for x in (enc_ids, enc_len, q_tok, q_len, q_terms))
    return lambda: eng.search_batch(*batch, out_scores_host=host_s, out_ids_host=host_i)


def main() -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--workload", choices=["search", "dashboard", "gpu"], default="search")
    ap.add_argument("--duration", type=float, default=300.0)
    ap.add_argument("--interval", type=float, default=5.0)
    ap.add_argument("--tolerance", type=float, default=0.10)
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    tracemalloc.start()
    with tempfile.TemporaryDirectory(prefix="infomesh-diag-") as tmp:
        step = {"search": _search_workload, "dashboard": _dashboard_workload, "gpu": _gpu_workload}[a.workload](Path(tmp))
        for _ in range(20):
            step()                                       # warm caches, lazy imports, CUDA graphs
        gc.collect()
        samples, calls, t0, nxt = [], 0, time.monotonic(), 0.0
        while (now := time.monotonic() - t0) < a.duration:
            step()
            calls += 1
            if now >= nxt:
                nxt += a.interval
                gc.collect()
                s = _sample(a.workload == "gpu")
                samples.append(s)
                if a.verbose:
                    print(f"[{now:7.1f}s] calls={calls} " + " ".join(f"{k}={v:.1f}" for k, v in s.items()), flush=True)
    if len(samples) < 8:
        print("too few samples; raise --duration or lower --interval")
        return 2
    q = len(samples) // 4
    bad = []
    print(f"{calls} calls in {a.duration:.0f} s ({calls / a.duration:.1f}/s), {len(samples)} samples")
    for k in samples[0]:
        first, last = statistics.median(s[k] for s in samples[:q]), statistics.median(s[k] for s in samples[-q:])
        growth = (last - first) / first if first > 0 else 0.0
        flag = growth > a.tolerance
        bad += [k] if flag else []
        print(f"  {k:18s} first-quarter {first:10.1f}  last-quarter {last:10.1f}  {growth:+7.1%} {'LEAK?' if flag else 'ok'}")
    return 1 if bad else 0


if __name__ == "__main__":
    raise SystemExit(main())
