"""Stand-in for the `chromadb` wheel (absent from this image) so the reference's UNMODIFIED ``VectorStore`` runs.

Covers the calls ``infomesh/index/vector_store.py`` makes: ``EphemeralClient`` / ``PersistentClient`` ->
``get_or_create_collection(name, metadata)`` -> ``upsert / delete / count / query``.  ChromaDB answers ``query`` with
hnswlib (approximate, CPU); this stand-in answers with an EXACT cosine search done by ``torch`` (library matmul + topk)
on the GPU when one is visible, CPU otherwise -- i.e. it is at least as fast and as accurate as the wheel it replaces,
so the reference arm is not handicapped.  No kernel, model or engine of infomesh_b200 is used."""
from __future__ import annotations

from typing import Any

import torch

__version__ = "0.0-shim"


class Collection:
    def __init__(self, name: str, metadata: dict | None = None):
        self.name = name
        self.metadata = metadata or {}
        self._dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self._emb: torch.Tensor | None = None     # [cap, dim] unit rows
        self._n = 0
        self._ids: list[str] = []
        self._row: dict[str, int] = {}
        self._meta: list[Any] = []
        self._docs: list[Any] = []
        self._alive: torch.Tensor | None = None

    def count(self) -> int:
        return len(self._row)

    def _reserve(self, extra: int, dim: int) -> None:
        need = self._n + extra
        if self._emb is None:
            cap = max(1024, need)
            self._emb = torch.zeros((cap, dim), device=self._dev, dtype=torch.float32)
            self._alive = torch.zeros((cap,), device=self._dev, dtype=torch.bool)
        elif need > self._emb.shape[0]:
            cap = max(need, self._emb.shape[0] * 2)
            e = torch.zeros((cap, dim), device=self._dev, dtype=torch.float32)
            e[:self._n] = self._emb[:self._n]
            a = torch.zeros((cap,), device=self._dev, dtype=torch.bool)
            a[:self._n] = self._alive[:self._n]
            self._emb, self._alive = e, a

    def upsert(self, ids, embeddings=None, metadatas=None, documents=None, **_k) -> None:
        e = torch.as_tensor(embeddings, dtype=torch.float32)
        if e.dim() == 1:
            e = e[None]
        e = torch.nn.functional.normalize(e.to(self._dev), dim=1)
        ids = [str(i) for i in ids]
        fresh = [j for j, i in enumerate(ids) if i not in self._row]
        self._reserve(len(fresh), e.shape[1])
        if len(fresh) == len(ids):              # bulk append
            a = self._n
            self._emb[a:a + len(ids)] = e
            self._alive[a:a + len(ids)] = True
            for j, i in enumerate(ids):
                self._row[i] = a + j
            self._ids.extend(ids)
            self._meta.extend(metadatas if metadatas is not None else [None] * len(ids))
            self._docs.extend(documents if documents is not None else [None] * len(ids))
            self._n += len(ids)
            return
        for j, i in enumerate(ids):
            r = self._row.get(i)
            if r is None:
                r = self._n
                self._n += 1
                self._row[i] = r
                self._ids.append(i)
                self._meta.append(None)
                self._docs.append(None)
            self._emb[r] = e[j]
            self._alive[r] = True
            if metadatas is not None:
                self._meta[r] = metadatas[j]
            if documents is not None:
                self._docs[r] = documents[j]

    add = upsert

    def delete(self, ids=None, **_k) -> None:
        for i in ids or []:
            r = self._row.pop(str(i), None)
            if r is not None:
                self._alive[r] = False

    def query(self, query_embeddings=None, n_results: int = 10, include=None, **_k) -> dict:
        out = {"ids": [], "distances": [], "metadatas": [], "documents": None, "embeddings": None}
        if self._emb is None or not self._row:
            return out
        q = torch.nn.functional.normalize(torch.as_tensor(query_embeddings, dtype=torch.float32).to(self._dev), dim=1)
        sim = q @ self._emb[:self._n].t()
        if len(self._row) != self._n:
            sim = sim.masked_fill(~self._alive[:self._n][None], -2.0)
        k = min(int(n_results), len(self._row))
        v, idx = torch.topk(sim, k, dim=1)
        v, idx = v.cpu().tolist(), idx.cpu().tolist()
        for row_v, row_i in zip(v, idx):
            out["ids"].append([self._ids[r] for r in row_i])
            out["distances"].append([1.0 - s for s in row_v])
            out["metadatas"].append([self._meta[r] or {} for r in row_i])
        return out


class _Client:
    def __init__(self, path: str | None = None, **_k):
        self._path = path
        self._cols: dict[str, Collection] = {}

    def get_or_create_collection(self, name: str, metadata: dict | None = None, **_k) -> Collection:
        if name not in self._cols:
            self._cols[name] = Collection(name, metadata)
        return self._cols[name]

    create_collection = get_or_create_collection
    get_collection = get_or_create_collection

    def delete_collection(self, name: str) -> None:
        self._cols.pop(name, None)

    def heartbeat(self) -> int:
        return 1


ClientAPI = _Client


def EphemeralClient(*a, **k) -> _Client:
    return _Client()


def PersistentClient(path: str | None = None, *a, **k) -> _Client:
    return _Client(path)


Client = EphemeralClient
