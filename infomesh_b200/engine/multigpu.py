"""Document-sharded serving over every GPU of a node: ``[gpu] devices = N`` (SURVEY §5 "scale-up inside a node").

The reference serves one process per node and scales by adding peers (infomesh/services.py:70-140 builds one store, one
vector index); a B200 node has 8 GPUs behind one NVSwitch, so the product path here is ONE front object with the
``GpuSearchIndex`` interface and one worker process per GPU:

* every worker opens the same ``LocalStore`` read-only, tokenises the WHOLE store once (global vocabulary, document
  frequencies and average length -- BM25 scores are then identical to a single-GPU index), keeps the postings, vectors
  and passages of its own contiguous document range ``[rank * per, (rank + 1) * per)`` in HBM, and joins an NCCL group
  for the bootstrap of the symmetric heaps;
* a search is a collective: the front hands the same query batch to every worker, each runs the fused pipeline
  (local BM25 + dense top-k pushed into every peer's heap, merge, RRF, passages pulled from the owning GPU, data-parallel
  cross-encoder, logit all-gather, final selection) and returns host arrays; rank 0's scores/rows are the answer, every
  rank contributes doc ids and K11 passage spans for the rows it owns;
* the front formats hits through its own store handle (``gpu_index.format_hits``).

Control plane: one duplex pipe per worker (``multiprocessing`` spawn context -- CUDA-safe); requests are fanned out before
any reply is read so the ranks enter the collective together; a worker that dies or misses its deadline marks the front
unhealthy and searches return empty until ``rebuild()`` respawns the group (the caller -- services / MCP -- falls back to
the CPU store exactly as it does when no GPU index exists).

``index_factory`` ("package.module:callable") replaces ``GpuSearchIndex`` inside the workers; the CPU test-suite uses it
to exercise the process plumbing and the merge without a GPU."""
from __future__ import annotations

import importlib
import multiprocessing as mp
import os
import socket
import threading
import time
from typing import Any

import numpy as np

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

READY_TIMEOUT_S = 1800.0
CALL_TIMEOUT_S = 120.0
AUTO_DOCS_PER_GPU = 250_000


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _resolve(path: str):
    mod, _, name = path.partition(":")
    return getattr(importlib.import_module(mod), name)


def _default_factory(store_path: str, rank: int, world: int, kwargs: dict):
    from infomesh_b200.engine.gpu_index import GpuSearchIndex
    from infomesh_b200.index.local_store import LocalStore
    from infomesh_b200.parallel import dist as D

    D.init()
    store = LocalStore(store_path)
    return GpuSearchIndex(store, device=f"cuda:{rank}", shard=(rank, world), **kwargs)


def _worker_main(rank: int, world: int, port: int, conn, store_path: str, kwargs: dict, factory: str | None) -> None:
    """Entry point of one per-GPU worker process."""
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    try:
        make = _resolve(factory) if factory else _default_factory
        kwargs = dict(kwargs)
        segments = kwargs.pop("__segments__", None)       # cold start from per-shard segment files instead of a rebuild
        index = make(store_path, rank, world, kwargs)
        n = index.load(shard_dir(segments, rank, world)) if segments else index.rebuild()
        conn.send(("ready", {"rank": rank, "documents": int(n), "stats": _plain(index.stats())}))
    except Exception as exc:  # noqa: BLE001 -- the front must learn why the worker never became ready
        conn.send(("failed", f"{type(exc).__name__}: {exc}"))
        return
    while True:
        try:
            op, arg = conn.recv()
        except (EOFError, OSError):
            break
        try:
            if op == "search":
                conn.send(("ok", index.search_arrays(arg)))
            elif op == "rebuild":
                conn.send(("ok", int(index.rebuild())))
            elif op == "stats":
                conn.send(("ok", _plain(index.stats())))
            elif op == "save":
                man = index.save(shard_dir(arg, rank, world))
                conn.send(("ok", {"rank": rank, "n_docs": int(man["n_docs"]), "model": man.get("model", ""),
                                  "bytes": int(sum(f["bytes"] for f in man["files"].values()))}))
            elif op == "load":
                conn.send(("ok", int(index.load(shard_dir(arg, rank, world)))))
            elif op == "delete":
                conn.send(("ok", bool(index.mark_deleted(int(arg)))))
            elif op == "close":
                conn.send(("ok", None))
                break
            else:
                conn.send(("error", f"unknown op {op!r}"))
        except Exception as exc:  # noqa: BLE001
            conn.send(("error", f"{type(exc).__name__}: {exc}"))
    closer = getattr(index, "close", None)
    if closer:
        closer()
    try:
        from infomesh_b200.parallel import dist as D

        D.shutdown()
    except Exception:  # noqa: BLE001
        pass


def shard_dir(root, rank: int, world: int):
    """Where shard ``rank`` of ``world`` keeps its segment files under ``root``."""
    from pathlib import Path

    return Path(root) / f"shard-{rank:02d}-of-{world:02d}"


def _plain(obj: Any) -> Any:
    """Stats dictionaries cross a pipe: keep them to builtin types."""
    if isinstance(obj, dict):
        return {str(k): _plain(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [_plain(v) for v in obj]
    if isinstance(obj, (np.integer,)):
        return int(obj)
    if isinstance(obj, (np.floating,)):
        return float(obj)
    return obj if isinstance(obj, (str, int, float, bool, type(None))) else str(obj)


class MultiGpuSearchIndex:
    """``GpuSearchIndex`` interface (``rebuild`` / ``search`` / ``search_many`` / ``stats`` / ``close``) over N worker
    processes, one per GPU."""

    def __init__(self, store: Any, *, devices: int, store_path: str | None = None, query_batch: int = 64,
                 index_factory: str | None = None, call_timeout: float = CALL_TIMEOUT_S, ready_timeout: float = READY_TIMEOUT_S,
                 **index_kwargs):
        if devices < 2:
            raise ValueError("MultiGpuSearchIndex needs devices >= 2 (use GpuSearchIndex for one GPU)")
        if query_batch % devices:
            raise ValueError(f"query_batch {query_batch} must be a multiple of devices {devices} (data-parallel reranker)")
        self.store, self.world, self.nq = store, int(devices), int(query_batch)
        self.store_path = str(store_path or getattr(store, "path", None) or getattr(store, "db_path", ""))
        self._factory, self._kwargs = index_factory, dict(index_kwargs, query_batch=query_batch)
        self._call_timeout, self._ready_timeout = float(call_timeout), float(ready_timeout)
        self._procs: list[mp.Process] = []
        self._conns: list[Any] = []
        self._lock = threading.Lock()           # one collective in flight: the pipes are not multiplexed
        self._worker_stats: list[dict] = []
        self.healthy = False
        self.built_at = 0.0
        self.build_seconds = 0.0
        self.reranker = None            # models live in the workers; host-side callers (mcp/handlers._reranker) must not reach for them
        self.device = "cuda"

    # ------------------------------------------------------------------ lifecycle
    def _spawn(self, segments: str | None = None) -> None:
        ctx = mp.get_context("spawn")
        port = _free_port()
        kwargs = dict(self._kwargs, __segments__=str(segments)) if segments else self._kwargs
        for rank in range(self.world):
            parent, child = ctx.Pipe(duplex=True)
            p = ctx.Process(target=_worker_main, name=f"infomesh-gpu{rank}", daemon=True,
                            args=(rank, self.world, port, child, self.store_path, kwargs, self._factory))
            p.start()
            child.close()
            self._procs.append(p)
            self._conns.append(parent)

    def _collect(self, timeout: float) -> list[Any]:
        """One reply per worker, in rank order, all within ``timeout`` seconds of now."""
        deadline = time.monotonic() + timeout
        out = []
        for rank, conn in enumerate(self._conns):
            left = deadline - time.monotonic()
            if left <= 0 or not conn.poll(left):
                alive = self._procs[rank].is_alive()
                raise TimeoutError(f"GPU worker {rank} {'did not answer in time' if alive else 'exited'}")
            tag, payload = conn.recv()
            if tag in ("failed", "error"):
                raise RuntimeError(f"GPU worker {rank}: {payload}")
            out.append(payload)
        return out

    def rebuild(self) -> int:
        """(Re)spawn the worker group if needed and rebuild every shard.  Returns the total document count."""
        t0 = time.time()
        with self._lock:
            try:
                if self.healthy:
                    for conn in self._conns:
                        conn.send(("rebuild", None))
                    counts = self._collect(self._ready_timeout)
                    n = int(sum(counts))
                else:
                    self._teardown()
                    self._spawn()
                    ready = self._collect(self._ready_timeout)
                    self._worker_stats = [r["stats"] for r in ready]
                    n = int(sum(r["documents"] for r in ready))
                self.healthy = True
            except Exception as exc:  # noqa: BLE001
                logger.error("multigpu_rebuild_failed", error=str(exc))
                self._teardown()
                raise
        self.n_docs, self._pending = n, 0
        self.built_at, self.build_seconds = time.time(), time.time() - t0
        logger.info("multigpu_index_built", docs=n, gpus=self.world, seconds=round(self.build_seconds, 2))
        return n

    # ------------------------------------------------------------------ persistence
    def save(self, directory) -> dict:
        """Every worker writes its shard's segment files under ``directory/shard-RR-of-WW`` (``GpuSearchIndex.save``); the
        front adds ``manifest.json`` naming the shards.  A node restarts from these with file reads + ``cudaMemcpy`` per
        GPU instead of re-tokenising and re-encoding the corpus on every rank."""
        import json
        from pathlib import Path

        if not self.healthy:
            raise RuntimeError("nothing to save: call rebuild() first")
        root = Path(directory)
        root.mkdir(parents=True, exist_ok=True)
        with self._lock:
            for conn in self._conns:
                conn.send(("save", str(root)))
            shards = self._collect(self._ready_timeout)
        manifest = {"format": 1, "world": self.world, "n_docs": int(sum(s["n_docs"] for s in shards)), "saved_at": time.time(),
                    "shards": [dict(s, dir=shard_dir("", s["rank"], self.world).name) for s in shards]}
        (root / "manifest.json").write_text(json.dumps(manifest, indent=1))
        return manifest

    def load(self, directory) -> int:
        """Inverse of :meth:`save`: (re)start the worker group from segment files written for the SAME number of GPUs.
        Returns the total document count; raises when the layout does not match or a shard refuses its segments."""
        import json
        from pathlib import Path

        root = Path(directory)
        man = json.loads((root / "manifest.json").read_text())
        if man.get("format") != 1 or int(man.get("world", 0)) != self.world:
            raise ValueError(f"segments were written for {man.get('world')} GPUs, this index runs on {self.world}")
        t0 = time.time()
        with self._lock:
            try:
                if self.healthy:
                    for conn in self._conns:
                        conn.send(("load", str(root)))
                    n = int(sum(self._collect(self._ready_timeout)))
                else:
                    self._teardown()
                    self._spawn(segments=str(root))
                    ready = self._collect(self._ready_timeout)
                    self._worker_stats = [r["stats"] for r in ready]
                    n = int(sum(r["documents"] for r in ready))
                self.healthy = True
            except Exception as exc:  # noqa: BLE001
                logger.error("multigpu_load_failed", error=str(exc))
                self._teardown()
                raise
        self.n_docs, self._pending = n, 0
        self.built_at, self.build_seconds = time.time(), time.time() - t0
        logger.info("multigpu_index_loaded", docs=n, gpus=self.world, seconds=round(self.build_seconds, 2))
        return n

    def refresh(self) -> int:
        """Sharded ranges are re-balanced on growth, so an append is a rebuild of every shard (in place, workers kept)."""
        before = getattr(self, "n_docs", 0)
        return max(0, self.rebuild() - before)

    def note_added(self, n: int = 1) -> None:
        self._pending = getattr(self, "_pending", 0) + n

    def _teardown(self) -> None:
        for conn in self._conns:
            try:
                conn.send(("close", None))
            except Exception:  # noqa: BLE001
                pass
        for p in self._procs:
            p.join(timeout=10)
            if p.is_alive():
                p.terminate()       # exact child handle, never a pattern
                p.join(timeout=5)
        for conn in self._conns:
            conn.close()
        self._procs, self._conns, self.healthy = [], [], False

    def close(self) -> None:
        with self._lock:
            self._teardown()

    @property
    def engine(self):
        """Truthy while the worker group serves (the serving surfaces test ``index.engine is not None``)."""
        return self if self.healthy else None

    def mark_deleted(self, doc_id: int) -> bool:
        """Tombstone a document on whichever shard owns it."""
        if not self.healthy:
            return False
        with self._lock:
            try:
                for conn in self._conns:
                    conn.send(("delete", int(doc_id)))
                return any(self._collect(min(self._call_timeout, 10.0)))
            except Exception as exc:  # noqa: BLE001
                logger.warning("multigpu_delete_failed", error=str(exc))
                return False

    # ------------------------------------------------------------------ queries
    def search_many(self, queries: list[str], k: int = 10) -> list[list[dict[str, object]]]:
        from infomesh_b200.engine.gpu_index import format_hits, merge_shard_arrays

        if not self.healthy or not queries:
            return [[] for _ in queries]
        out: list[list[dict[str, object]]] = []
        for a in range(0, len(queries), self.nq):
            chunk = queries[a:a + self.nq]
            with self._lock:
                if not self.healthy:
                    out.extend([] for _ in chunk)
                    continue
                try:
                    for conn in self._conns:          # fan out first: the ranks meet inside the fused exchange
                        conn.send(("search", chunk))
                    parts = self._collect(self._call_timeout)
                except Exception as exc:  # noqa: BLE001
                    logger.error("multigpu_search_failed", error=str(exc))
                    self._teardown()                  # a rank is gone: peers would spin on its flags forever
                    out.extend([] for _ in chunk)
                    continue
            out.extend(format_hits(self.store, chunk, k, merge_shard_arrays(parts)) if parts and parts[0] else [[] for _ in chunk])
        return out

    def search(self, query: str, k: int = 10) -> list[dict[str, object]]:
        return self.search_many([query], k)[0]

    def stats(self) -> dict[str, object]:
        per_rank = self._worker_stats
        if self.healthy:
            with self._lock:
                try:
                    for conn in self._conns:
                        conn.send(("stats", None))
                    per_rank = self._worker_stats = self._collect(min(self._call_timeout, 10.0))
                except Exception as exc:  # noqa: BLE001
                    logger.warning("multigpu_stats_failed", error=str(exc))
        return {"gpus": self.world, "healthy": self.healthy, "documents": getattr(self, "n_docs", 0), "built_at": self.built_at,
                "build_seconds": round(self.build_seconds, 2), "query_batch": self.nq,
                "hbm_bytes": int(sum(int(s.get("hbm_bytes", 0)) for s in per_rank)), "per_rank": per_rank}


def make_index(store: Any, gcfg: Any = None, **overrides):
    """The serving surfaces' single entry point: ``[gpu] devices`` > 1 (or 0 = all visible, when more than one GPU is
    visible and the store is large enough to shard) -> :class:`MultiGpuSearchIndex`, else ``GpuSearchIndex``."""
    from infomesh_b200.engine.gpu_index import GpuSearchIndex, gpu_index_kwargs

    kw = dict(gpu_index_kwargs(gcfg), **overrides)
    want = int(getattr(gcfg, "devices", 1) or 0) if gcfg is not None else 1
    if want != 1:
        import torch

        visible = torch.cuda.device_count() if torch.cuda.is_available() else 0
        n_docs = int(store.get_stats().get("document_count", 0)) if hasattr(store, "get_stats") else 0
        # auto (0): one GPU per AUTO_DOCS_PER_GPU documents -- a small store gains nothing from 8 copies of the models
        n = min(visible, max(1, n_docs // AUTO_DOCS_PER_GPU)) if want == 0 else min(want, visible)
        nq = int(kw.get("query_batch", getattr(gcfg, "query_batch", 64)))
        while n > 1 and (nq % n or n_docs < 2 * n):
            n -= 1
        if n > 1:
            kw.setdefault("query_batch", nq)
            return MultiGpuSearchIndex(store, devices=n, **kw)
    dev = int(getattr(gcfg, "device", 0)) if gcfg is not None else 0
    return GpuSearchIndex(store, device=f"cuda:{dev}", **kw)


def warm_start(index: Any, store: Any, segments_dir: str | os.PathLike | None) -> int:
    """Bring ``index`` (single- or multi-GPU) up: from the segment files under ``segments_dir`` when they describe the
    store as it is now, else by a full build whose result is written back there.  Returns the document count.

    "As it is now" = same document count and same highest document id as the store reports; anything else (documents
    added or removed since the save, another GPU count, another encoder, a checksum failure) falls through to a rebuild --
    stale segments are never served."""
    import json
    from pathlib import Path

    if not segments_dir:
        return int(index.rebuild())
    root = Path(segments_dir)
    stats = store.get_stats() if hasattr(store, "get_stats") else {}
    want = int(stats.get("document_count", -1))
    try:
        man = json.loads((root / "manifest.json").read_text())
        if want >= 0 and int(man.get("n_docs", -2)) == want and _segments_cover(man, root, store):
            n = int(index.load(root))
            logger.info("gpu_index_warm_start", docs=n, source=str(root))
            return n
    except FileNotFoundError:
        pass
    except Exception as exc:  # noqa: BLE001 -- unreadable / mismatching segments: rebuild
        logger.warning("gpu_segments_rejected", error=str(exc))
    n = int(index.rebuild())
    if n:
        try:
            index.save(root)
        except Exception as exc:  # noqa: BLE001 -- a full disk must not take the freshly built index down
            logger.warning("gpu_segments_save_failed", error=str(exc))
    return n


def _segments_cover(man: dict, root, store: Any) -> bool:
    """Does the highest document id in the segments equal the store's?  (Single-GPU manifests carry ``doc_id_range``; a
    sharded save carries it per shard.)"""
    import json

    top = -1
    if "doc_id_range" in man:
        top = int(man["doc_id_range"][1])
    for sh in man.get("shards", ()):
        sub = root / sh["dir"] / "manifest.json"
        if sub.exists():
            top = max(top, int(json.loads(sub.read_text()).get("doc_id_range", [0, -1])[1]))
    newest = getattr(store, "max_doc_id", None)
    return top < 0 or newest is None or int(newest()) == top
