"""HBM-resident mirror of a :class:`LocalStore`, searchable through the fused hybrid pipeline.

This is the bridge between the CPU plane (SQLite documents, the reference's ``search_hybrid`` API surface,
infomesh/search/query.py:244-319) and the device plane (``engine.hybrid.HybridEngine``).  A rebuild streams the
store once and produces, on the GPU:

* dense vectors  — encoder forward over ``title + text`` in batches (replaces ChromaDB/hnswlib, N2/N3);
* BM25 postings  — C++ ``IndexBuilder`` tokenises on the host, CSR arrays are uploaded once (replaces FTS5 scoring, N1);
* passage tokens — the leading ``passage_len`` reranker tokens of every document for pair assembly;
* passage terms  — every document split into passages (paragraph / sentence rules of ``search.passage.split_passages``)
  and each passage's BM25 term ids, so the best passage of every hit is chosen ON THE DEVICE by ``passage_score``
  (K11: coverage + 0.1 x density, the reference's ``select_best_passage``, infomesh/search/passage.py:143-227) and only
  its character span is cut out of the document on the host.

Queries are answered in batches: ``search_many`` tokenises on the host, copies three small int tensors from pinned
memory, replays the captured CUDA graph and maps the winning doc rows back to URLs / titles / snippets through SQLite.
Deleted documents are masked with the ``alive`` byte map until the next rebuild.

Two guards for serving:

* **No ranking by untrained models.**  The dense leg is used only when the encoder carries checkpoint weights
  (``BertModel.pretrained``, set by :mod:`infomesh_b200.models.loader`) and the cross-encoder second stage only when the
  reranker does; otherwise the device path answers with BM25 order (GPU postings) alone.  Benchmarks and kernel tests
  pass ``allow_untrained=True``.
* **Thread safety.**  MCP / HTTP call ``search_many`` from worker threads while the crawl loop may ``rebuild``: one
  re-entrant lock serialises device passes over the shared pinned staging buffers, and a rebuild prepares everything
  off to the side and publishes engine + row maps + staging in one step under that lock.
"""
from __future__ import annotations

import threading
import time
from dataclasses import dataclass
from typing import Any

import numpy as np
import torch

from infomesh_b200.engine.hybrid import HybridConfig, HybridEngine
from infomesh_b200.models.bert import BGE_RERANKER_BASE, BGE_SMALL, BertModel
from infomesh_b200.ops.bm25 import Bm25Index, HostIndexBuilder
from infomesh_b200.utils.log import get_logger
from infomesh_b200.utils.tokenizer import BERT_SPECIALS, XLMR_SPECIALS, HashTokenizer

logger = get_logger(__name__)

LATENCY_NQ = 8          # queries per pass of the low-latency lane (interactive, one-at-a-time searches)
UNKNOWN_TERM = -2       # out-of-vocabulary query term: makes the implicit AND empty (FTS5 semantics)


@dataclass
class _ShardCfg:
    doc_base: int = 0


class _StoreShard:
    """The attribute set HybridEngine expects from a shard."""

    def __init__(self, device, vectors, bm25, passage_tok, passage_len, alive):
        self.device, self.vectors, self.bm25 = device, vectors, bm25
        self.passage_tok, self.passage_len, self.alive = passage_tok, passage_len, alive
        self.neg_age = self.authority = None
        self.cfg = _ShardCfg(0)

    def nbytes(self) -> int:
        return self.vectors.numel() * 2 + self.bm25.nbytes() + self.passage_tok.numel() * 4 + self.passage_len.numel() * 4


class GpuSearchIndex:
    def __init__(self, store: Any, *, device: str | torch.device = "cuda:0", encoder: BertModel | None = None,
                 reranker: BertModel | None = None, rerank: bool = True, query_batch: int = 64, passage_len: int = 96,
                 enc_doc_tokens: int = 128, embed_batch: int = 256, use_graph: bool = True, seed: int = 0,
                 encoder_path: str | None = None, reranker_path: str | None = None, allow_untrained: bool = False,
                 rank_signals: bool = False, authority_fn=None, shard: tuple[int, int] | None = None, dense_dtype: str = "bf16"):
        self.store = store
        self.device = torch.device(device)
        self.enc_tok = self.rr_tok = None
        if encoder is None and encoder_path:
            from infomesh_b200.models.loader import load_bert, load_tokenizer

            encoder = load_bert(encoder_path, device=self.device)
            self.enc_tok = load_tokenizer(encoder_path, encoder.cfg.vocab_size)
        if reranker is None and reranker_path and rerank:
            from infomesh_b200.models.loader import load_bert, load_tokenizer

            reranker = load_bert(reranker_path, device=self.device)
            self.rr_tok = load_tokenizer(reranker_path, reranker.cfg.vocab_size)
        self.encoder = encoder or BertModel(BGE_SMALL, device=self.device, seed=seed + 1)
        self.allow_untrained = bool(allow_untrained)
        # BM25-only answers ordered like the reference's search_local (freshness / trust / authority fused on the device)
        self.rank_signals, self.authority_fn = bool(rank_signals), authority_fn
        # (rank, world) of a document-sharded deployment (engine/multigpu.py): this process indexes documents
        # [rank * per, (rank + 1) * per) of the store on its GPU; vocabulary / df / idf stay global
        self.dense_dtype = dense_dtype if dense_dtype in ("bf16", "fp8") else "bf16"     # [gpu] shard_dtype
        self.shard_rank, self.shard_world = shard if shard is not None else (0, 1)
        self.row_base = 0
        self._pheap = None
        # ranking policy: a model without checkpoint weights may not influence results (see module docstring)
        self.use_dense = self.allow_untrained or bool(getattr(self.encoder, "pretrained", False))
        want_rr = rerank and (self.allow_untrained or bool(getattr(reranker, "pretrained", False)))
        self.reranker = (reranker or BertModel(BGE_RERANKER_BASE, device=self.device, seed=seed + 2)) if want_rr else None
        if rerank and not want_rr:
            logger.warning("gpu_rerank_disabled", reason="reranker has no checkpoint weights (random init); BM25/RRF order is kept")
        if not self.use_dense:
            logger.warning("gpu_dense_disabled", reason="encoder has no checkpoint weights (random init); BM25-only retrieval")
        self.enc_tok = self.enc_tok or HashTokenizer(self.encoder.cfg.vocab_size, BERT_SPECIALS)
        self.rr_tok = self.rr_tok or HashTokenizer(BGE_RERANKER_BASE.vocab_size, XLMR_SPECIALS)
        self.rerank, self.nq, self.passage_len = want_rr, query_batch, passage_len
        self.enc_doc_tokens, self.embed_batch, self.use_graph = enc_doc_tokens, embed_batch, use_graph
        self.builder: HostIndexBuilder | None = None
        self.engine: HybridEngine | None = None
        self.doc_ids = np.zeros(0, dtype=np.int64)      # row -> LocalStore doc_id
        self._row_of: dict[int, int] = {}
        self._pending = 0
        self.built_at = 0.0
        self.build_seconds = 0.0
        self._lock = threading.RLock()

    # ------------------------------------------------------------------ build
    @property
    def n_docs(self) -> int:
        return int(self.doc_ids.size)

    def rebuild(self) -> int:
        """Stream the store and (re)create every device structure.  Returns the number of documents."""
        t0 = time.time()
        dev = self.device
        builder = HostIndexBuilder()
        ids: list[int] = []
        vec_chunks: list[torch.Tensor] = []
        pt_rows: list[list[int]] = []
        batch_text: list[str] = []
        docs_text: list[str] = []
        crawled: list[float] = []
        auth: list[float] = []

        def flush():
            if batch_text:
                tok, lens = self.enc_tok.encode_batch(batch_text, max_len=self.enc_doc_tokens)
                vec_chunks.append(self.encoder.embed(tok.to(dev, non_blocking=True), lens.to(dev, non_blocking=True)))
                batch_text.clear()

        n_total = int(self.store.get_stats().get("document_count", 0)) if self.shard_world > 1 else 0
        per = (n_total + self.shard_world - 1) // self.shard_world if self.shard_world > 1 else 0
        lo, hi = (self.shard_rank * per, min(n_total, (self.shard_rank + 1) * per)) if self.shard_world > 1 else (0, 1 << 62)
        seen = -1
        for doc in self.store.iter_documents():
            body = f"{doc.title}\n{doc.text}"
            builder.add_text(body)                            # EVERY rank tokenises the whole store: global vocabulary and df
            seen += 1
            # ranking signals are per-document scalars: kept for EVERY document on every rank (8 B / doc), because the
            # rank fuse runs after the exchange, on candidate rows of any shard
            crawled.append(float(doc.crawled_at or 0.0))
            if self.authority_fn is not None:
                try:
                    auth.append(float(self.authority_fn(doc.url)))
                except Exception:  # noqa: BLE001
                    auth.append(0.0)
            if not (lo <= seen < hi):
                continue
            ids.append(int(doc.doc_id))
            pt_rows.append(self.rr_tok.encode_plain(body, self.passage_len))
            docs_text.append(doc.text)
            batch_text.append(body[:2000])                   # the reference embeds the first 2000 chars (vector_store.py:156)
            if len(batch_text) >= self.embed_batch:
                flush()
        flush()
        n = len(ids)
        if n == 0:
            with self._lock:
                self.doc_ids, self._row_of, self._pending = np.zeros(0, dtype=np.int64), {}, 0
                self.engine, self.builder = None, builder
            return 0
        csr = builder.export()
        if self.shard_world > 1:
            csr = _slice_csr(csr, lo, hi)                       # this rank's postings, global df / avg_len kept alongside
            csr["per_rank"] = per
            self.row_base = lo
        passages = _passage_arrays(builder, docs_text)          # after every term is registered
        t_ref = t0
        if self.shard_world > 1:      # one reference instant for every rank: the fused scores must agree bit for bit
            import torch.distributed as dist

            box = torch.tensor([t0], dtype=torch.float64, device=dev)
            dist.broadcast(box, src=0)
            t_ref = float(box.item())
        self._t_ref = t_ref
        passages["neg_age"] = np.asarray(crawled, dtype=np.float64) - t_ref   # seconds before this build (<= 0), fp32-safe
        passages["authority"] = np.asarray(auth, dtype=np.float32) if auth else None
        del docs_text
        vectors = torch.cat(vec_chunks).contiguous()
        ptok = torch.full((n, self.passage_len), self.rr_tok.sp.pad, dtype=torch.int32)
        plen = torch.zeros((n,), dtype=torch.int32)
        for i, row in enumerate(pt_rows):
            if row:
                ptok[i, :len(row)] = torch.tensor(row, dtype=torch.int32)
            plen[i] = max(len(row), 1)
        self._builder_live = True
        self._install(builder, csr, vectors, ptok.to(dev), plen.to(dev), torch.ones((n,), dtype=torch.uint8, device=dev),
                      np.asarray(ids, dtype=np.int64), passages)
        shard = self.engine.shard
        self.built_at, self.build_seconds = time.time(), time.time() - t0
        logger.info("gpu_index_built", docs=n, seconds=round(self.build_seconds, 2), hbm_mb=round(shard.nbytes() / 2 ** 20, 1))
        return n

    def refresh(self) -> int:
        """Incremental append: index the documents added to the store since the last build WITHOUT re-reading, re-tokenising
        or re-encoding the ones already resident.  New rows are tokenised into the live posting builder, encoded, and
        concatenated behind the resident vectors / passage tokens / passage tables; the CSR is re-exported from the builder
        (a memcpy-class pass over the postings) and the device structures are swapped in atomically.  Cost is
        O(new documents) model work + O(postings) copies, against O(corpus) model work for :meth:`rebuild`.
        Returns the number of documents appended.  (A sharded index re-balances its ranges, so it rebuilds.)"""
        if (self.engine is None or self.builder is None or self.shard_world > 1 or self.doc_ids.size == 0
                or not getattr(self, "_builder_live", True)):      # loaded segments: the builder holds the vocabulary only
            before = self.n_docs
            return max(0, self.rebuild() - before)
        t0 = time.time()
        dev = self.device
        last = int(self.doc_ids.max())
        new_ids, texts, bodies, pt_rows, crawled, auth = [], [], [], [], [], []
        for doc in self.store.iter_documents(after=last):
            body = f"{doc.title}\n{doc.text}"
            self.builder.add_text(body)
            new_ids.append(int(doc.doc_id))
            texts.append(doc.text)
            bodies.append(body[:2000])
            pt_rows.append(self.rr_tok.encode_plain(body, self.passage_len))
            crawled.append(float(doc.crawled_at or 0.0))
            if self.authority_fn is not None:
                try:
                    auth.append(float(self.authority_fn(doc.url)))
                except Exception:  # noqa: BLE001
                    auth.append(0.0)
        m = len(new_ids)
        if m == 0:
            self._pending = 0
            return 0
        old = self.engine.shard
        vec_new = []
        for a in range(0, m, self.embed_batch):
            tok, lens = self.enc_tok.encode_batch(bodies[a:a + self.embed_batch], max_len=self.enc_doc_tokens)
            vec_new.append(self.encoder.embed(tok.to(dev, non_blocking=True), lens.to(dev, non_blocking=True)))
        vectors = torch.cat([old.vectors, *vec_new]).contiguous()
        ptok_new = torch.full((m, self.passage_len), self.rr_tok.sp.pad, dtype=torch.int32)
        plen_new = torch.zeros((m,), dtype=torch.int32)
        for i, row in enumerate(pt_rows):
            if row:
                ptok_new[i, :len(row)] = torch.tensor(row, dtype=torch.int32)
            plen_new[i] = max(len(row), 1)
        ptok = torch.cat([old.passage_tok, ptok_new.to(dev)])
        plen = torch.cat([old.passage_len, plen_new.to(dev)])
        alive = torch.cat([old.alive, torch.ones((m,), dtype=torch.uint8, device=dev)])       # deletions survive an append
        fresh = _passage_arrays(self.builder, texts)
        prev = self._pass
        n_tok_prev = int(prev["off"][-1])
        passages = {"terms": np.concatenate([prev["terms"][:n_tok_prev], fresh["terms"]]),
                    "off": np.concatenate([prev["off"], fresh["off"][1:] + n_tok_prev]),
                    "doc_off": np.concatenate([prev["doc_off"], fresh["doc_off"][1:] + prev["doc_off"][-1]]),
                    "span": np.concatenate([prev["span"].reshape(-1, 2), fresh["span"]])}
        if prev.get("neg_age") is not None:
            passages["neg_age"] = np.concatenate([prev["neg_age"], np.asarray(crawled, dtype=np.float64) - getattr(self, "_t_ref", t0)])
            passages["authority"] = (np.concatenate([prev["authority"], np.asarray(auth, dtype=np.float32)])
                                     if prev.get("authority") is not None and auth else prev.get("authority"))
        csr = self.builder.export()
        self._install(self.builder, csr, vectors, ptok, plen, alive, np.concatenate([self.doc_ids, np.asarray(new_ids, dtype=np.int64)]), passages)
        self.built_at = time.time()
        logger.info("gpu_index_appended", docs=m, total=self.n_docs, seconds=round(time.time() - t0, 2))
        return m

    # ------------------------------------------------------------------ persistence
    _FILES = ("vectors.bin", "csr_off.bin", "csr_doc.bin", "csr_tf.bin", "doc_len.bin", "df.bin", "passage_tok.bin",
              "passage_len.bin", "doc_ids.bin", "alive.bin", "vocab.txt", "pass_terms.bin", "pass_off.bin", "doc_pass_off.bin",
              "pass_span.bin", "neg_age.bin", "authority.bin")

    def _model_tag(self) -> str:
        """Identifies the encoder whose vectors are stored: config + a checksum of its first projection matrix."""
        import hashlib

        w = self.encoder.w.layers[0]["wqkv"][:8].float().cpu().numpy().tobytes()
        return f"{self.encoder.cfg.name}:{self.encoder.cfg.hidden}x{self.encoder.cfg.layers}:{hashlib.sha256(w).hexdigest()[:16]}"

    def save(self, directory) -> dict:
        """Write the device structures as flat binary segments + ``manifest.json`` (sizes, sha256, model tag) so a node
        cold-starts with file reads and ``cudaMemcpy`` instead of re-encoding and re-tokenising the corpus."""
        import hashlib
        import json
        from pathlib import Path

        if self.engine is None:
            raise RuntimeError("nothing to save: call rebuild() first")
        d = Path(directory)
        d.mkdir(parents=True, exist_ok=True)
        sh, bm = self.engine.shard, self.engine.shard.bm25
        csr = self._csr
        arrays = {"vectors.bin": sh.vectors.view(torch.int16).cpu().numpy(), "csr_off.bin": csr["off"], "csr_doc.bin": csr["doc"],
                  "csr_tf.bin": csr["tf"], "doc_len.bin": csr["doc_len"], "df.bin": csr["df"],
                  "passage_tok.bin": sh.passage_tok.cpu().numpy(), "passage_len.bin": sh.passage_len.cpu().numpy(),
                  "doc_ids.bin": self.doc_ids, "alive.bin": sh.alive.cpu().numpy(), "pass_terms.bin": self._pass["terms"],
                  "pass_off.bin": self._pass["off"], "doc_pass_off.bin": self._pass["doc_off"], "pass_span.bin": self._pass["span"]}
        # ranking signals (age at build time, authority) and the instant they are relative to: a loaded index must rank
        # exactly like the one that was saved
        if self._pass.get("neg_age") is not None:
            arrays["neg_age.bin"] = np.asarray(self._pass["neg_age"], dtype=np.float64)
        if self._pass.get("authority") is not None:
            arrays["authority.bin"] = np.asarray(self._pass["authority"], dtype=np.float32)
        files = {}
        for name, arr in arrays.items():
            arr = np.ascontiguousarray(arr)
            arr.tofile(d / name)
            files[name] = {"bytes": int(arr.nbytes), "dtype": str(arr.dtype), "sha256": hashlib.sha256(arr.tobytes()).hexdigest()}
        vocab = "\n".join(self.builder.term(t) for t in range(self.builder.vocab))
        (d / "vocab.txt").write_text(vocab, encoding="utf-8")
        files["vocab.txt"] = {"bytes": len(vocab.encode()), "dtype": "utf-8", "sha256": hashlib.sha256(vocab.encode()).hexdigest()}
        manifest = {"format": 1, "n_docs": self.n_docs, "dim": int(sh.vectors.shape[1]), "passage_len": self.passage_len, "vocab": self.builder.vocab,
                    "avg_len": bm.avg_len, "model": self._model_tag(), "doc_id_range": [int(self.doc_ids.min()), int(self.doc_ids.max())],
                    "saved_at": time.time(), "t_ref": float(getattr(self, "_t_ref", 0.0)), "files": files}
        if self.shard_world > 1:
            # one directory per shard (engine/multigpu.py names them): the global BM25 statistics ride in the manifest, the
            # global document frequencies are df.bin itself (the slice keeps the whole-corpus df)
            manifest["shard"] = {"rank": self.shard_rank, "world": self.shard_world, "row_base": int(self.row_base),
                                 "per_rank": int(csr["per_rank"]), "n_docs_global": int(csr["n_docs_global"]),
                                 "avg_len_global": float(csr["avg_len_global"])}
        (d / "manifest.json").write_text(json.dumps(manifest, indent=1))
        return manifest

    def load(self, directory, *, verify: bool = True) -> int:
        """Inverse of :meth:`save`.  Refuses segments written for a different encoder or with a failing checksum."""
        import hashlib
        import json
        from pathlib import Path

        d = Path(directory)
        man = json.loads((d / "manifest.json").read_text())
        if man.get("format") != 1:
            raise ValueError("unknown segment format")
        if man["model"] != self._model_tag():
            raise ValueError(f"segments were built for encoder {man['model']}, this index uses {self._model_tag()}")
        t0 = time.time()

        def read(name, dtype):
            raw = np.fromfile(d / name, dtype=dtype)
            meta = man["files"][name]
            if raw.nbytes != meta["bytes"] or (verify and hashlib.sha256(raw.tobytes()).hexdigest() != meta["sha256"]):
                raise ValueError(f"segment {name} is corrupt")
            return raw

        n, dim, dev = man["n_docs"], man["dim"], self.device
        csr = {"off": read("csr_off.bin", np.int64), "doc": read("csr_doc.bin", np.int32), "tf": read("csr_tf.bin", np.uint8),
               "doc_len": read("doc_len.bin", np.int32), "df": read("df.bin", np.int32)}
        vectors = torch.from_numpy(read("vectors.bin", np.int16)).view(torch.bfloat16).view(n, dim).to(dev)
        ptok = torch.from_numpy(read("passage_tok.bin", np.int32)).view(n, man["passage_len"]).to(dev)
        plen = torch.from_numpy(read("passage_len.bin", np.int32)).to(dev)
        alive = torch.from_numpy(read("alive.bin", np.uint8)).to(dev)
        doc_ids = read("doc_ids.bin", np.int64)
        passages = {"terms": read("pass_terms.bin", np.int32), "off": read("pass_off.bin", np.int64),
                    "doc_off": read("doc_pass_off.bin", np.int64), "span": read("pass_span.bin", np.int32).reshape(-1, 2)}
        if "neg_age.bin" in man["files"]:
            passages["neg_age"] = read("neg_age.bin", np.float64)
            passages["authority"] = read("authority.bin", np.float32) if "authority.bin" in man["files"] else None
            self._t_ref = float(man.get("t_ref", 0.0))
        sh_meta = man.get("shard")
        if (sh_meta is None) != (self.shard_world == 1) or (sh_meta and (sh_meta["rank"], sh_meta["world"]) != (self.shard_rank, self.shard_world)):
            raise ValueError(f"segments are for shard {sh_meta and (sh_meta['rank'], sh_meta['world'])}, "
                             f"this index is shard {(self.shard_rank, self.shard_world)}")
        if sh_meta:
            csr.update(df_global=csr["df"], n_docs_global=sh_meta["n_docs_global"], avg_len_global=sh_meta["avg_len_global"],
                       per_rank=sh_meta["per_rank"])
            self.row_base = int(sh_meta["row_base"])
        builder = HostIndexBuilder()
        terms = (d / "vocab.txt").read_text(encoding="utf-8")
        if terms:
            builder.tokenize(terms.replace("\n", " "), add=True)       # re-registers the terms in id order
        if builder.vocab != man["vocab"]:
            raise ValueError("vocabulary does not round-trip")
        self.passage_len = man["passage_len"]
        self._builder_live = False
        self._install(builder, csr, vectors, ptok, plen, alive, doc_ids, passages)
        self.built_at, self.build_seconds = time.time(), time.time() - t0
        return n

    def _install(self, builder, csr, vectors, ptok, plen, alive, doc_ids, passages) -> None:
        """Adopt device structures (fresh build or loaded segments): the engine, row maps and pinned staging are built
        into locals and published together under the lock, so a concurrent search sees either the old or the new index."""
        dev = self.device
        bm = Bm25Index(csr, device=dev, n_docs_global=csr.get("n_docs_global"), avg_len=csr.get("avg_len_global"),
                       df_global=csr.get("df_global"))
        ptabs, ekw = None, {}
        if self.shard_world > 1:
            # passage tokens live in a symmetric heap so pair assembly on any rank can read this shard's passages over NVLink
            from infomesh_b200.parallel import dist as D
            from infomesh_b200.parallel import symm

            per = int(csr["per_rank"])
            if self._pheap is not None:
                self._pheap.close()
            self._pheap = symm.SymmetricHeap(per * (self.passage_len + 1) * 4 + (4 << 20), D.ctx())
            tok_v, tok_off = self._pheap.alloc((per, self.passage_len), torch.int32)
            len_v, len_off = self._pheap.alloc((per,), torch.int32)
            tok_v[:ptok.shape[0]].copy_(ptok)
            len_v[:plen.shape[0]].copy_(plen)
            ptok, plen = tok_v[:ptok.shape[0]], len_v[:plen.shape[0]]
            ptabs = (self._pheap.peer_table(tok_off), self._pheap.peer_table(len_off))
            self._pheap.barrier()
            ekw = dict(passage_tables=ptabs, docs_per_shard=per, passage_len=self.passage_len)
        shard = _StoreShard(dev, vectors, bm, ptok, plen, alive)
        shard.cfg = _ShardCfg(self.row_base)
        if passages.get("neg_age") is not None:
            shard.neg_age = torch.from_numpy(np.minimum(passages["neg_age"], 0.0).astype(np.float32)).to(dev)
            shard.authority = torch.from_numpy(passages["authority"]).to(dev) if passages.get("authority") is not None else None
            shard.signal_base = 0 if self.shard_world > 1 else self.row_base     # sharded: the signal arrays cover every document
        cfg = HybridConfig(nq=self.nq, rerank=self.rerank, k_fetch=20, n_rerank=20, k_out=10, pair_seq=min(128, 32 + self.passage_len),
                           use_graph=self.use_graph, dense=self.use_dense, rank_signals=self.rank_signals,
                           dense_dtype=self.dense_dtype if vectors.shape[1] % 128 == 0 else "bf16",
                           degraded_ok=self.shard_world > 1)       # sharded serving answers without a silent / unhealthy shard
        pin = torch.cuda.is_available()

        def lane(nq: int):
            """One serving lane: an engine captured for ``nq`` queries per pass + its pinned staging buffers."""
            from dataclasses import replace as _r

            eng = HybridEngine(shard, _r(cfg, nq=nq), encoder=self.encoder, reranker=self.reranker, **ekw)
            eng.warm()          # graph capture belongs to the build, not to the first query (and never to a serving thread)

            def mk(*shape, fill=0):
                t = torch.full(shape, fill, dtype=torch.int32)
                return t.pin_memory() if pin else t

            h_scores = torch.empty((nq, cfg.k_out), dtype=torch.float32)
            h_ids = torch.empty((nq, cfg.k_out), dtype=torch.int64)
            if pin:
                h_scores, h_ids = h_scores.pin_memory(), h_ids.pin_memory()
            return eng, (mk(nq, cfg.enc_seq), mk(nq, fill=1), mk(nq, cfg.max_q_tokens, fill=self.rr_tok.sp.pad), mk(nq, fill=1),
                         mk(nq, cfg.max_terms, fill=-1), h_scores, h_ids)

        engine, staging = lane(cfg.nq)
        # interactive queries arrive one at a time: a pass captured for LATENCY_NQ queries costs a fraction of a full-batch
        # pass (the cross-encoder sees 8 x 20 pairs instead of 64 x 20), so single searches take the small lane
        small = lane(LATENCY_NQ) if cfg.nq > LATENCY_NQ and LATENCY_NQ % max(self.shard_world, 1) == 0 else None
        row_of = {int(x): i for i, x in enumerate(doc_ids)}
        pass_dev = {k: torch.from_numpy(np.ascontiguousarray(passages[k])).to(dev) for k in ("terms", "off", "doc_off")}
        with self._lock:
            self._pass, self._pass_dev = passages, pass_dev
            self._csr, self.engine, self.builder = csr, engine, builder
            self._small = small
            self.doc_ids, self._row_of, self._pending = doc_ids, row_of, 0
            (self._h_enc, self._h_enc_len, self._h_qtok, self._h_qlen, self._h_terms, self._h_scores, self._h_ids) = staging

    def mark_deleted(self, doc_id: int) -> bool:
        with self._lock:
            row = self._row_of.get(int(doc_id))
            if row is None or self.engine is None:
                return False
            self.engine.shard.alive[row] = 0
            return True

    def note_added(self, n: int = 1) -> None:
        """New documents become searchable on the GPU at the next rebuild; callers rebuild past a threshold."""
        self._pending += n

    @property
    def stale(self) -> bool:
        return self._pending > 0

    # ------------------------------------------------------------------ query
    def _stage(self, queries: list[str], engine=None, staging=None) -> None:
        engine = engine or self.engine
        h_enc, h_enc_len, h_qtok, h_qlen, h_terms = (staging or (self._h_enc, self._h_enc_len, self._h_qtok, self._h_qlen, self._h_terms))[:5]
        cfg = engine.cfg
        h_enc.zero_()
        h_enc_len.fill_(2)
        h_qtok.fill_(self.rr_tok.sp.pad)
        h_qlen.fill_(1)
        h_terms.fill_(-1)
        for i, q in enumerate(queries):
            e = self.enc_tok.encode(q, cfg.enc_seq)
            h_enc[i, :len(e)] = torch.tensor(e, dtype=torch.int32)
            h_enc_len[i] = len(e)
            t = self.rr_tok.encode_plain(q, cfg.max_q_tokens) or [self.rr_tok.sp.unk]
            h_qtok[i, :len(t)] = torch.tensor(t, dtype=torch.int32)
            h_qlen[i] = len(t)
            terms = [int(x) if x >= 0 else UNKNOWN_TERM for x in dict.fromkeys(self.builder.tokenize(q).tolist())][:cfg.max_terms]
            if terms:
                h_terms[i, :len(terms)] = torch.tensor(terms, dtype=torch.int32)
        for i in range(len(queries), cfg.nq):        # padding rows: a CLS/SEP-only query with no terms
            e = [self.enc_tok.sp.cls, self.enc_tok.sp.sep]
            h_enc[i, :2] = torch.tensor(e, dtype=torch.int32)

    def search_arrays(self, chunk: list[str]) -> dict[str, np.ndarray]:
        """One device pass over up to ``query_batch`` queries -> host arrays ``[nq, k_out]``: ``scores``, global ``rows``,
        and for the rows THIS process owns ``doc_ids`` (else -1) and the best passage ``pass`` / its character ``span``.
        In a sharded deployment every rank runs this collectively on the same queries (engine/multigpu.py merges)."""
        with self._lock:      # staging buffers, graph-static device buffers and the row map are shared state
            if self.engine is None:
                return {}
            self._searches = getattr(self, "_searches", 0) + 1
            if self.shard_world > 1 and self._searches % 256 == 1:       # ECC / Xid state -> degraded-mode mask of the exchange
                from infomesh_b200.resources.gpu_health import GpuHealthMonitor

                mon = self.__dict__.setdefault("_mesh_health", GpuHealthMonitor(None))
                try:
                    self.engine.apply_health(mon)
                except Exception:  # noqa: BLE001 -- health polling must never fail a search
                    pass
            small = getattr(self, "_small", None)
            eng, st = small if (small is not None and len(chunk) <= small[0].cfg.nq) else (
                self.engine, (self._h_enc, self._h_enc_len, self._h_qtok, self._h_qlen, self._h_terms, self._h_scores, self._h_ids))
            self._stage(chunk, eng, st)
            eng.search_batch(*st)
            # K11 on the device: best passage (coverage + 0.1 x density) of every returned (query, document) pair
            best_pass = self._best_passages(eng)
            torch.cuda.current_stream(self.device).synchronize()
            scores, rows = st[5].numpy()[:len(chunk)].copy(), st[6].numpy()[:len(chunk)].copy()
            best_pass = best_pass.cpu().numpy().reshape(eng.cfg.nq, -1)[:len(chunk)]
            doc_ids, spans, doc_off = self.doc_ids, self._pass["span"], self._pass["doc_off"]
        local = rows - self.row_base
        own = (local >= 0) & (local < doc_ids.size)
        lr = np.where(own, local, 0)
        ids = np.where(own, doc_ids[lr] if doc_ids.size else -1, -1)
        pidx = doc_off[lr] + np.maximum(best_pass, 0)
        has = own & (best_pass >= 0) & (pidx < doc_off[np.minimum(lr + 1, doc_off.size - 1)])
        span = np.full(rows.shape + (2,), -1, np.int64)
        if spans.shape[0]:
            span[has] = spans[pidx[has]]
        return {"scores": scores, "rows": rows, "doc_ids": ids.astype(np.int64), "pass": np.where(has, best_pass, -1), "span": span}

    def search_many(self, queries: list[str], k: int = 10) -> list[list[dict[str, object]]]:
        """One device pass per ``query_batch`` queries.  Each hit: doc_id, url, title, snippet, score."""
        if self.engine is None or not queries:
            return [[] for _ in queries]
        out: list[list[dict[str, object]]] = []
        nq = self.engine.cfg.nq
        for a in range(0, len(queries), nq):
            chunk = queries[a:a + nq]
            arr = self.search_arrays(chunk)
            out.extend(format_hits(self.store, chunk, k, arr) if arr else [[] for _ in chunk])
        return out

    def _best_passages(self, eng=None) -> torch.Tensor:
        """int32 ``[nq * k_out]``: index of the best passage inside each returned document (-1: none / no query terms)."""
        from infomesh_b200.ops.bm25 import passage_score

        eng = eng or self.engine
        nq, k_out = eng.out_ids.shape
        pair_doc = eng.out_ids.reshape(-1) - self.row_base          # rows of other shards are not scored here
        pair_doc = torch.where((pair_doc >= 0) & (pair_doc < self.doc_ids.size), pair_doc, -1).to(torch.int32)
        pair_query = torch.arange(nq, device=self.device, dtype=torch.int32).repeat_interleave(k_out)
        pd = self._pass_dev
        _s, best = passage_score(pd["terms"], pd["off"], pd["doc_off"], pair_doc, pair_query, eng.in_terms)
        return best

    def search(self, query: str, k: int = 10) -> list[dict[str, object]]:
        return self.search_many([query], k)[0]

    def close(self) -> None:
        """Drop the device structures (and, sharded, leave the symmetric heaps: peers must do the same collectively)."""
        with self._lock:
            eng, self.engine = self.engine, None
            small, self._small = getattr(self, "_small", None), None
            for e in (eng, small[0] if small else None):
                if e is not None and getattr(e, "heap", None) is not None:
                    e.heap.close()
            if self._pheap is not None:
                self._pheap.close()
                self._pheap = None

    def health(self) -> dict:
        """NVML view of the serving device(s): ECC / Xid / throttle state (resources/gpu_health.py); cached monitor."""
        from infomesh_b200.resources.gpu_health import GpuHealthMonitor

        mon = self.__dict__.get("_health_mon")
        if mon is None:
            mon = self.__dict__["_health_mon"] = GpuHealthMonitor([self.device.index or 0] if self.device.type == "cuda" else None)
        return mon.summary()

    def stats(self) -> dict[str, object]:
        sh = self.engine.shard if self.engine else None
        return {"health": self.health(),"documents": self.n_docs, "pending": self._pending, "built_at": self.built_at, "build_seconds": round(self.build_seconds, 2),
                "hbm_bytes": sh.nbytes() if sh else 0, "vocab": self.builder.vocab if self.builder else 0,
                "query_batch": self.nq, "rerank": self.rerank, "dense": self.use_dense,
                "encoder": getattr(self.encoder, "source", "random-init"),
                "reranker": getattr(self.reranker, "source", "random-init") if self.reranker is not None else "off",
                "cuda_graph": bool(self.engine and self.engine._graph is not None)}


def format_hits(store, chunk: list[str], k: int, arr: dict[str, np.ndarray]) -> list[list[dict[str, object]]]:
    """Host-side tail of a search: document rows from the store, snippet = the device-chosen passage (K11) or a window
    around the first query term.  ``arr`` is :meth:`GpuSearchIndex.search_arrays` output (merged over shards)."""
    out = []
    for i, q in enumerate(chunk):
        hits: list[dict[str, object]] = []
        for j in range(arr["scores"].shape[1]):
            did = int(arr["doc_ids"][i, j])
            if did < 0 or len(hits) >= k:
                continue
            doc = store.get_document(did)
            if doc is None:
                continue
            a, b = (int(x) for x in arr["span"][i, j])
            snippet = doc.text[a:b][:300] if a >= 0 else _snippet(doc.text, q)
            hits.append({"doc_id": doc.doc_id, "url": doc.url, "title": doc.title, "snippet": snippet,
                         "score": float(arr["scores"][i, j]), "crawled_at": doc.crawled_at, "passage": int(arr["pass"][i, j])})
        out.append(hits)
    return out


def merge_shard_arrays(parts: list[dict[str, np.ndarray]]) -> dict[str, np.ndarray]:
    """Combine the per-rank views of ONE collective search.  Rank 0's scores / rows are the answer; ``doc_ids`` / ``pass`` /
    ``span`` of a row come from the rank that owns it, matched BY ROW VALUE inside the same query -- the final selection
    runs on every rank and near-ties may land in a different slot there, so positions are not trusted."""
    out = {key: np.array(val, copy=True) for key, val in parts[0].items()}
    want = out["rows"]                                                     # [nq, k]
    for part in parts[1:]:
        have = part["rows"]
        same = (have[:, :, None] == want[:, None, :]) & (part["doc_ids"][:, :, None] >= 0) & (want[:, None, :] >= 0)   # [nq, k_src, k_dst]
        src = same.argmax(axis=1)                                          # for each destination slot: source slot (if any)
        hit = same.any(axis=1) & (out["doc_ids"] < 0)
        q = np.broadcast_to(np.arange(want.shape[0])[:, None], want.shape)
        for key in ("doc_ids", "pass", "span"):
            out[key][hit] = part[key][q[hit], src[hit]]
    return out


def _slice_csr(csr: dict, lo: int, hi: int) -> dict:
    """Postings of documents ``[lo, hi)`` out of a global CSR, re-based to local rows, with the GLOBAL statistics BM25
    needs (document count, average length, document frequencies) carried along."""
    off, doc, tf, dl, df = csr["off"], csr["doc"], csr["tf"], csr["doc_len"], csr["df"]
    V = len(off) - 1
    nnz = int(off[-1])
    term_of = np.repeat(np.arange(V, dtype=np.int64), np.diff(off))
    d = doc[:nnz]
    sel = (d >= lo) & (d < hi)
    counts = np.bincount(term_of[sel], minlength=V)
    new_off = np.zeros(V + 1, np.int64)
    new_off[1:] = np.cumsum(counts)
    n_all = len(dl)
    return {"off": new_off, "doc": (d[sel] - lo).astype(np.int32), "tf": tf[:nnz][sel], "doc_len": dl[lo:hi], "df": df,
            "df_global": df, "n_docs_global": n_all, "avg_len_global": float(dl.sum() / max(n_all, 1)),
            "per_rank": 0}


def _passage_arrays(builder: HostIndexBuilder, texts: list[str]) -> dict:
    """Split every document into passages and tokenise them into BM25 term ids.

    -> ``terms`` int32 (all passages back to back), ``off`` int64 ``[n_pass + 1]`` token offsets, ``doc_off`` int64
    ``[n_docs + 1]`` first passage of each document, ``span`` int32 ``[n_pass, 2]`` character range of the passage in the
    document text (what the host cuts out once the device has picked the passage)."""
    from infomesh_b200.search.passage import split_passages

    terms: list[np.ndarray] = []
    off, doc_off, span = [0], [0], []
    for text in texts:
        cursor = 0
        for p in split_passages(text):
            head = p[:40]
            at = text.find(head, cursor)
            if at < 0:
                at = cursor
            end = min(len(text), at + len(p))
            cursor = max(cursor, end - 1)
            t = builder.tokenize(text[at:end])
            t = t[t >= 0]
            terms.append(t.astype(np.int32))
            off.append(off[-1] + int(t.size))
            span.append((at, end))
        doc_off.append(len(span))
    return {"terms": np.concatenate(terms) if terms else np.zeros(1, np.int32), "off": np.asarray(off, np.int64),
            "doc_off": np.asarray(doc_off, np.int64), "span": np.asarray(span, np.int32).reshape(-1, 2)}


def gpu_index_kwargs(gcfg) -> dict:
    """``[gpu]`` config section -> keyword arguments of :class:`GpuSearchIndex` (checkpoint paths + ranking policy)."""
    kw = {}
    if gcfg is None:
        return kw
    if getattr(gcfg, "encoder_path", ""):
        kw["encoder_path"] = gcfg.encoder_path
    if getattr(gcfg, "reranker_path", ""):
        kw["reranker_path"] = gcfg.reranker_path
    kw["rank_signals"] = True            # serving surfaces answer like search_local: ranking signals fused on the device
    kw["dense_dtype"] = str(getattr(gcfg, "shard_dtype", "bf16"))
    kw["allow_untrained"] = bool(getattr(gcfg, "allow_untrained_models", False))
    return kw


def _snippet(text: str, query: str, width: int = 200) -> str:
    """Window around the first query-term occurrence (cheap host-side stand-in for FTS5 snippet())."""
    low = text.lower()
    pos = min((p for p in (low.find(t) for t in query.lower().split()) if p >= 0), default=0)
    a = max(0, pos - width // 4)
    s = text[a:a + width].strip()
    return ("…" if a > 0 else "") + s + ("…" if a + width < len(text) else "")
