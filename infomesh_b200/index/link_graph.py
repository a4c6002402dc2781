"""Link graph + domain authority (domain-level PageRank) in SQLite.

Reference infomesh/index/link_graph.py:19-307: damping 0.85, <= 20 iterations, L1 convergence 1e-6, same-domain
links weighted x0.1, scores normalised by the maximum; ``url_authority`` feeds ``rank_local_results``.
The power iteration runs as a dense/sparse matrix-vector product (NumPy here; ``spmv_pagerank`` on GPU is K13).
"""
from __future__ import annotations

import sqlite3
import threading
from pathlib import Path
from urllib.parse import urlparse

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_DAMPING = 0.85
_MAX_ITERATIONS = 20
_CONVERGENCE_THRESHOLD = 1e-6
_SAME_DOMAIN_WEIGHT = 0.1


def _cuda_available() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


class LinkGraph:
    def __init__(self, db_path: str | None = None):
        self._path = str(db_path) if db_path else ":memory:"
        if self._path != ":memory:":
            Path(self._path).parent.mkdir(parents=True, exist_ok=True)
        self._conn = sqlite3.connect(self._path, check_same_thread=False)
        self._conn.row_factory = sqlite3.Row
        self._conn.execute("PRAGMA journal_mode=WAL")
        self._conn.execute("PRAGMA busy_timeout=5000")
        self._lock = threading.RLock()
        self._conn.executescript("""
            CREATE TABLE IF NOT EXISTS links (
                source_url TEXT NOT NULL, target_url TEXT NOT NULL,
                source_domain TEXT NOT NULL, target_domain TEXT NOT NULL,
                discovered_at REAL NOT NULL DEFAULT (strftime('%s', 'now')),
                PRIMARY KEY (source_url, target_url));
            CREATE INDEX IF NOT EXISTS idx_links_target_domain ON links(target_domain);
            CREATE INDEX IF NOT EXISTS idx_links_source_domain ON links(source_domain);
            CREATE TABLE IF NOT EXISTS domain_authority (
                domain TEXT PRIMARY KEY, score REAL NOT NULL DEFAULT 0.0,
                inbound_count INTEGER NOT NULL DEFAULT 0, outbound_count INTEGER NOT NULL DEFAULT 0,
                updated_at REAL NOT NULL DEFAULT (strftime('%s', 'now')));
        """)
        self._conn.commit()

    @staticmethod
    def _extract_domain(url: str) -> str:
        try:
            return (urlparse(url).netloc or "").lower()
        except ValueError:
            return ""

    def add_links(self, source_url: str, target_urls: list[str]) -> int:
        """Record outbound links of a crawled page; returns how many new edges were stored."""
        src = self._extract_domain(source_url)
        if not src:
            return 0
        rows = []
        for t in target_urls:
            dom = self._extract_domain(t)
            if dom and t != source_url:
                rows.append((source_url, t, src, dom))
        if not rows:
            return 0
        with self._lock:
            before = self._conn.total_changes
            self._conn.executemany("INSERT OR IGNORE INTO links (source_url, target_url, source_domain, target_domain) "
                                   "VALUES (?, ?, ?, ?)", rows)
            self._conn.commit()
            return self._conn.total_changes - before

    def compute_domain_authority(self, use_gpu: bool | None = None) -> dict[str, float]:
        """Power iteration over the domain graph.  ``use_gpu`` (default: automatically for >= 50k domains on a CUDA box)
        runs the iteration with the edge-parallel kernel of ``ops/graph.py`` instead of NumPy."""
        import numpy as np

        edges = self._conn.execute("SELECT source_domain s, target_domain t, COUNT(*) c FROM links "
                                   "GROUP BY source_domain, target_domain").fetchall()
        if not edges:
            return {}
        domains = sorted({e["s"] for e in edges} | {e["t"] for e in edges})
        idx = {d: i for i, d in enumerate(domains)}
        n = len(domains)
        src = np.fromiter((idx[e["s"]] for e in edges), dtype=np.int64, count=len(edges))
        dst = np.fromiter((idx[e["t"]] for e in edges), dtype=np.int64, count=len(edges))
        w = np.fromiter((float(e["c"]) * (_SAME_DOMAIN_WEIGHT if e["s"] == e["t"] else 1.0) for e in edges),
                        dtype=np.float64, count=len(edges))
        out_w = np.bincount(src, weights=w, minlength=n)
        share = np.divide(w, out_w[src], out=np.zeros_like(w), where=out_w[src] > 0)
        if use_gpu is None:
            use_gpu = n >= 50_000 and _cuda_available()
        score = None
        if use_gpu:
            try:
                from infomesh_b200.ops.graph import pagerank

                score = pagerank(src, dst, w, n, _DAMPING, _MAX_ITERATIONS, _CONVERGENCE_THRESHOLD).double().cpu().numpy()
            except Exception:  # noqa: BLE001 — fall back to the CPU loop
                logger.exception("authority_gpu_failed")
                score = None
        iterate = score is None
        if iterate:
            score = np.full(n, 1.0 / n)
        for it in range(_MAX_ITERATIONS if iterate else 0):
            nxt = np.full(n, (1.0 - _DAMPING) / n)
            np.add.at(nxt, dst, _DAMPING * score[src] * share)
            delta = float(np.abs(nxt - score).sum())
            score = nxt
            if delta < _CONVERGENCE_THRESHOLD:
                logger.debug("authority_converged", iterations=it + 1, diff=delta)
                break
        top = float(score.max())
        norm = score / top if top > 0 else score
        ext = src != dst
        inbound = {d: 0 for d in domains}
        outbound = {d: 0 for d in domains}
        for s_i, d_i in {(int(a), int(b)) for a, b in zip(src[ext], dst[ext])}:
            inbound[domains[d_i]] += 1
            outbound[domains[s_i]] += 1
        result = {d: float(norm[i]) for d, i in idx.items()}
        with self._lock:
            self._conn.executemany(
                "INSERT OR REPLACE INTO domain_authority (domain, score, inbound_count, outbound_count, updated_at) "
                "VALUES (?, ?, ?, ?, strftime('%s', 'now'))",
                [(d, round(result[d], 6), inbound[d], outbound[d]) for d in domains])
            self._conn.commit()
        logger.info("domain_authority_computed", domains=n)
        return result

    def domain_authority(self, domain: str) -> float:
        row = self._conn.execute("SELECT score FROM domain_authority WHERE domain = ?", (domain.lower(),)).fetchone()
        return float(row["score"]) if row else 0.0

    def url_authority(self, url: str) -> float:
        dom = self._extract_domain(url)
        return self.domain_authority(dom) if dom else 0.0

    def get_stats(self) -> dict[str, int]:
        links = self._conn.execute("SELECT COUNT(*) c FROM links").fetchone()["c"]
        doms = self._conn.execute("SELECT COUNT(*) c FROM domain_authority").fetchone()["c"]
        return {"link_count": int(links), "domain_count": int(doms)}

    def close(self) -> None:
        try:
            self._conn.close()
        except sqlite3.Error:
            pass

    def __enter__(self) -> "LinkGraph":
        return self

    def __exit__(self, *exc: object) -> None:
        self.close()
