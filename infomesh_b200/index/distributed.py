"""Keyword -> peer pointer index published through the DHT (reference infomesh/index/distributed.py:29-358).

The WAN analogue of the intra-node sharding: documents are partitioned over peers and a keyword lookup returns
pointers ``(peer_id, doc_id, url, score, title)``; queries sum scores per (peer, doc) across keywords.
"""
from __future__ import annotations

import re
from collections import Counter
from dataclasses import dataclass
from typing import Any

from infomesh_b200.p2p.protocol import PeerPointer
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MIN_KEYWORD_LENGTH = 2
MAX_KEYWORDS_PER_DOC = 50
MAX_POINTERS_PER_KEYWORD = 100
_WORD = re.compile(r"\b[a-zA-Z0-9]+\b")
_STOP_WORDS = frozenset(
    "a an the and or but in on at to for of is it be as do by he we so if no up my me am us are was has had not all "
    "can her his its our you who how did get may new now old see way from with this that have will been each make "
    "like than them then into over such when very what just also more some only come could would about which their "
    "there these those other after being where does".split())


def extract_keywords(text: str, *, max_keywords: int = MAX_KEYWORDS_PER_DOC) -> list[str]:
    """Most frequent alphanumeric words (len >= 2, stop words removed); ties keep first-seen order."""
    counts = Counter(w for w in _WORD.findall(text.lower()) if len(w) >= MIN_KEYWORD_LENGTH and w not in _STOP_WORDS)
    return [w for w, _ in sorted(counts.items(), key=lambda kv: kv[1], reverse=True)[:max_keywords]]


@dataclass
class DistributedIndexStats:
    documents_published: int = 0
    keywords_published: int = 0
    queries_performed: int = 0
    pointers_found: int = 0


def _as_str(v: object, default: str = "") -> str:
    return v if isinstance(v, str) else default


def _as_int(v: object, default: int = 0) -> int:
    if isinstance(v, bool):
        return default
    try:
        return int(v) if isinstance(v, (int, str)) else default
    except ValueError:
        return default


def _as_float(v: object, default: float = 0.0) -> float:
    if isinstance(v, bool):
        return default
    try:
        return float(v) if isinstance(v, (int, float, str)) else default
    except ValueError:
        return default


class DistributedIndex:
    def __init__(self, dht: Any, local_peer_id: str):
        self._dht = dht
        self._peer_id = local_peer_id
        self._stats = DistributedIndexStats()

    @property
    def stats(self) -> DistributedIndexStats:
        return self._stats

    async def publish_document(self, doc_id: int, url: str, title: str, text: str, score: float = 1.0) -> int:
        return await self.publish_batch([dict(doc_id=doc_id, url=url, title=title, text=text, score=score)])

    async def publish_batch(self, documents: list[dict[str, Any]]) -> int:
        """One DHT write per keyword for the whole batch (<= 100 pointers per keyword)."""
        by_keyword: dict[str, list[dict[str, object]]] = {}
        docs_used = 0
        for doc in documents:
            doc_id, url, text = _as_int(doc.get("doc_id")), _as_str(doc.get("url")), _as_str(doc.get("text"))
            if doc_id <= 0 or not url or not text:
                continue
            kws = extract_keywords(text)
            if not kws:
                continue
            docs_used += 1
            ptr = {"peer_id": self._peer_id, "doc_id": doc_id, "url": url,
                   "score": _as_float(doc.get("score"), 1.0), "title": _as_str(doc.get("title"))}
            for kw in kws:
                bucket = by_keyword.setdefault(kw, [])
                if len(bucket) < MAX_POINTERS_PER_KEYWORD:
                    bucket.append(ptr)
        published = 0
        for kw, ptrs in by_keyword.items():
            if await self._dht.publish_keyword(kw, ptrs):
                published += 1
        self._stats.documents_published += docs_used
        self._stats.keywords_published += published
        logger.debug("distributed_index_batch_published", documents=len(documents), used=docs_used,
                     keywords_published=published)
        return published

    async def query(self, keywords: list[str]) -> list[PeerPointer]:
        self._stats.queries_performed += 1
        agg: dict[tuple[str, int], dict[str, Any]] = {}
        for kw in keywords:
            for ptr in await self._dht.query_keyword(kw):
                if not isinstance(ptr, dict):
                    continue
                key = (_as_str(ptr.get("peer_id")), _as_int(ptr.get("doc_id")))
                if key in agg:
                    agg[key]["score"] = _as_float(agg[key].get("score")) + _as_float(ptr.get("score"), 0.5)
                else:
                    agg[key] = dict(ptr)
        ranked = sorted(agg.values(), key=lambda p: _as_float(p.get("score")), reverse=True)
        self._stats.pointers_found += len(ranked)
        return [PeerPointer(peer_id=_as_str(p.get("peer_id")), doc_id=_as_int(p.get("doc_id")),
                            url=_as_str(p.get("url")), score=_as_float(p.get("score")), title=_as_str(p.get("title")))
                for p in ranked]
