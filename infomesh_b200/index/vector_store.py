"""VectorStore — dense index + encoder with the reference's API, backed by the B200 engine.

Reference: ChromaDB (hnswlib ANN, cosine) + ``SentenceTransformer("all-MiniLM-L6-v2")``
(infomesh/index/vector_store.py:26-274).  Here the encoder is the hand-written BERT stack
(``models.bert``: bge-small-en / MiniLM-L6 shape, random-init unless weights are supplied) and search is the
**exact** fused similarity-GEMM + top-k kernel (``ops.search.sim_topk``) over vectors resident in HBM; on a machine
without CUDA everything falls back to PyTorch on the CPU so the API stays usable (config #1 plumbing).

Semantics kept: text to embed is ``f"{title}. {text}"[:2000]``, preview is ``text[:500]``, similarity =
cosine in [-1, 1] rounded to 4 dp, ``min_score`` filter, upsert by doc_id, persistence under ``persist_dir``.
"""
from __future__ import annotations

import json
import time
from dataclasses import dataclass
from pathlib import Path

import torch

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

DEFAULT_MODEL = "all-MiniLM-L6-v2"
COLLECTION_NAME = "infomesh_docs"
MAX_EMBED_CHARS = 2000
PREVIEW_CHARS = 500


@dataclass(frozen=True)
class VectorSearchResult:
    doc_id: str
    url: str
    title: str
    text_preview: str
    score: float


class VectorStore:
    def __init__(self, persist_dir: Path | str | None = None, model_name: str = DEFAULT_MODEL, collection_name: str = "infomesh_docs", *,
                 device: str | None = None, max_seq_len: int = 256, capacity: int = 4096, seed: int = 0):
        self._collection_name = collection_name          # kept for callers that name their collection; one collection per store here
        from infomesh_b200.models.bert import CONFIGS, BertModel
        from infomesh_b200.utils.tokenizer import BERT_SPECIALS, HashTokenizer

        self._persist = Path(persist_dir) if persist_dir else None
        self._model_name = model_name
        cfg = CONFIGS.get(model_name) or CONFIGS["bge-small-en"]
        self._device = torch.device(device or ("cuda" if torch.cuda.is_available() else "cpu"))
        self._on_gpu = self._device.type == "cuda"
        if self._on_gpu:
            from infomesh_b200 import _native

            _native.require()
        self._model = BertModel(cfg, device=self._device, seed=seed)
        self._tok = HashTokenizer(cfg.vocab_size, BERT_SPECIALS)
        self._max_seq = min(max_seq_len, cfg.max_pos)
        self._dim = cfg.hidden
        self._vecs = torch.zeros((capacity, self._dim), device=self._device, dtype=torch.bfloat16)
        self._alive = torch.zeros((capacity,), device=self._device, dtype=torch.uint8)
        self._n = 0
        self._meta: list[dict | None] = []
        self._slot_of: dict[str, int] = {}
        if self._persist is not None:
            self._persist.mkdir(parents=True, exist_ok=True)
            self._load()
        logger.debug("vector_store_initialized", model=cfg.name, device=str(self._device), docs=len(self._slot_of))

    # ------------------------------------------------------------------ embedding
    def _embed(self, texts: list[str]) -> torch.Tensor:
        """L2-normalised embeddings [len(texts), dim] (bf16 on GPU, fp32 on CPU)."""
        ids, lens = self._tok.encode_batch(texts, max_len=self._max_seq, pad_to_multiple=16)
        ids, lens = ids.to(self._device), lens.to(self._device)
        with torch.no_grad():
            if self._on_gpu:
                return self._model.embed(ids, lens)
            return self._model.embed_ref(ids, lens)

    def embed(self, texts: list[str]) -> list[list[float]]:
        return self._embed(texts).float().cpu().tolist()

    # ------------------------------------------------------------------ writes
    def _grow(self, need: int) -> None:
        cap = self._vecs.shape[0]
        if need <= cap:
            return
        new_cap = max(need, cap * 2)
        v = torch.zeros((new_cap, self._dim), device=self._device, dtype=torch.bfloat16)
        a = torch.zeros((new_cap,), device=self._device, dtype=torch.uint8)
        v[:cap], a[:cap] = self._vecs, self._alive
        self._vecs, self._alive = v, a

    def add_document(self, doc_id: int, url: str, title: str, text: str, *, language: str | None = None) -> None:
        self.add_documents([dict(doc_id=doc_id, url=url, title=title, text=text, language=language)])

    def add_documents(self, docs: list[dict]) -> None:
        """Batched upsert (one encoder pass for the whole batch)."""
        if not docs:
            return
        emb = self._embed([f"{d['title']}. {d['text']}"[:MAX_EMBED_CHARS] for d in docs]).to(torch.bfloat16)
        for row, d in zip(emb, docs):
            key = str(d["doc_id"])
            slot = self._slot_of.get(key)
            if slot is None:
                slot = self._n
                self._grow(slot + 1)
                self._n += 1
                self._meta.append(None)
                self._slot_of[key] = slot
            self._vecs[slot] = row
            self._alive[slot] = 1
            self._meta[slot] = {"doc_id": key, "url": d["url"], "title": d["title"],
                                "text_preview": d["text"][:PREVIEW_CHARS], "language": d.get("language") or ""}
        self._save()

    def delete_document(self, doc_id: int) -> None:
        slot = self._slot_of.pop(str(doc_id), None)
        if slot is not None:
            self._alive[slot] = 0
            self._vecs[slot] = 0
            self._meta[slot] = None
            self._save()

    # ------------------------------------------------------------------ search
    def search(self, query: str, *, limit: int = 10, min_score: float = 0.0) -> list[VectorSearchResult]:
        t0 = time.monotonic()
        live = len(self._slot_of)
        if live == 0 or not query:
            return []
        k = max(1, min(int(limit), live))
        q = self._embed([query])
        if self._on_gpu and k <= 32:
            from infomesh_b200.ops.search import sim_topk

            scores, ids = sim_topk(q.to(torch.bfloat16), self._vecs[:self._n], k, alive=self._alive[:self._n])
            scores, ids = scores[0].tolist(), ids[0].tolist()
        else:
            sc = (q.float() @ self._vecs[:self._n].float().t())[0]
            sc = sc.masked_fill(self._alive[:self._n] == 0, float("-inf"))
            top = torch.topk(sc, k)
            scores, ids = top.values.tolist(), top.indices.tolist()
        out: list[VectorSearchResult] = []
        for s, i in zip(scores, ids):
            if i < 0 or s == float("-inf"):
                continue
            meta = self._meta[i]
            if meta is None:
                continue
            sim = round(float(s), 4)
            if sim < min_score:
                continue
            out.append(VectorSearchResult(meta["doc_id"], meta["url"], meta["title"], meta["text_preview"], sim))
        logger.debug("vector_search", query=query[:60], results=len(out),
                     elapsed_ms=round((time.monotonic() - t0) * 1000, 1))
        return out

    # ------------------------------------------------------------------ persistence / stats
    def _save(self) -> None:
        if self._persist is None:
            return
        torch.save({"vecs": self._vecs[:self._n].cpu(), "alive": self._alive[:self._n].cpu()},
                   self._persist / "vectors.pt")
        (self._persist / "meta.json").write_text(json.dumps({"model": self._model_name, "meta": self._meta}), "utf-8")

    def _load(self) -> None:
        vp, mp = self._persist / "vectors.pt", self._persist / "meta.json"
        if not (vp.exists() and mp.exists()):
            return
        blob = torch.load(vp, map_location="cpu")
        meta = json.loads(mp.read_text("utf-8"))["meta"]
        n = len(meta)
        self._grow(n)
        self._vecs[:n] = blob["vecs"].to(self._device)
        self._alive[:n] = blob["alive"].to(self._device)
        self._n, self._meta = n, meta
        self._slot_of = {m["doc_id"]: i for i, m in enumerate(meta) if m is not None}

    def get_stats(self) -> dict[str, int | str]:
        return {"document_count": len(self._slot_of), "model": self._model_name, "collection": COLLECTION_NAME,
                "dimension": self._dim, "device": str(self._device)}

    def close(self) -> None:
        self._save()
