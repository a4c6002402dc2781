"""Second-stage re-ranking of the top <= 20 candidates.

Two judges behind one call shape:
  * :func:`rerank_with_llm` — the reference's prompt-an-LLM path (numbered list in, JSON permutation out, original
    order kept on any failure) (reference infomesh/search/reranker.py:20-163);
  * :func:`rerank_with_cross_encoder` — the B200 path: (query, title + snippet) pairs scored by the in-process
    cross-encoder (bge-reranker-base on the native kernels), one batched forward, argsort by logit.
"""
from __future__ import annotations

import json
import re
from typing import Any

from infomesh_b200.index.ranking import RankedResult
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MAX_RERANK_CANDIDATES = 20
_ARRAY = re.compile(r"\[[\d\s,]*\]")
_RERANK_PROMPT = ("You are a search result relevance judge. Given a search query and a list of search results, "
                  "re-rank them by relevance to the query.\n\nQuery: {query}\n\nResults:\n{results_block}\n\n"
                  "Return ONLY a JSON array of result numbers in order of relevance, most relevant first.\n"
                  "Example: [3, 1, 5, 2, 4]\n\nYour ranking (JSON array only):")


def _build_results_block(results: list[RankedResult], *, max_snippet: int = 150) -> str:
    return "\n".join(f"{i}. [{r.title}] " + r.snippet[:max_snippet].replace("\n", " ") for i, r in enumerate(results, 1))


def _parse_ranking_response(response: str, count: int) -> list[int] | None:
    """1-based JSON array -> 0-based permutation; junk entries dropped, unmentioned results appended in order."""
    m = _ARRAY.search(response)
    if not m:
        return None
    try:
        raw = json.loads(m.group())
    except ValueError:
        return None
    if not isinstance(raw, list):
        return None
    order: list[int] = []
    for v in raw:
        if isinstance(v, int) and not isinstance(v, bool) and 1 <= v <= count and v - 1 not in order:
            order.append(v - 1)
    return order + [i for i in range(count) if i not in order]


async def rerank_with_llm(query: str, results: list[RankedResult], llm_backend: Any, *, top_n: int | None = None,
                          max_candidates: int = MAX_RERANK_CANDIDATES) -> list[RankedResult]:
    from infomesh_b200.summarizer.engine import LLMBackend

    if not isinstance(llm_backend, LLMBackend) or not results:
        return results
    head, tail = results[:max_candidates], results[max_candidates:]
    try:
        if not await llm_backend.is_available():
            return results
        reply = await llm_backend.generate(_RERANK_PROMPT.format(query=query, results_block=_build_results_block(head)), max_tokens=256)
        order = _parse_ranking_response(reply, len(head))
    except Exception as exc:  # noqa: BLE001 — ranking must degrade to the first-stage order, never fail the search
        logger.warning("rerank_failed", error=str(exc))
        return results
    if order is None:
        return results
    out = [head[i] for i in order] + tail
    return out[:top_n] if top_n else out


class CrossEncoderReranker:
    """Holds the cross-encoder and its tokenizer; ``score(query, passages)`` is one batched GPU forward."""

    def __init__(self, model: Any | None = None, *, device: str | None = None, max_len: int = 256, max_query_tokens: int = 32):
        import torch

        from infomesh_b200.models.bert import BGE_RERANKER_BASE, BertModel
        from infomesh_b200.utils.tokenizer import XLMR_SPECIALS, HashTokenizer

        self.device = torch.device(device or ("cuda:0" if torch.cuda.is_available() else "cpu"))
        self.model = model or BertModel(BGE_RERANKER_BASE, device=self.device)
        self.tok = HashTokenizer(self.model.cfg.vocab_size, XLMR_SPECIALS)
        self.max_len, self.max_q = max_len, max_query_tokens

    def encode_pairs(self, query: str, passages: list[str]):
        """``<s> q </s></s> p </s>`` rows, padded to a multiple of 8 (the device pair-builder's layout)."""
        import torch

        sp = self.tok.sp
        q = self.tok.encode_plain(query, self.max_q)
        rows = []
        for p in passages:
            room = self.max_len - len(q) - 4
            rows.append([sp.cls, *q, sp.sep, sp.sep, *self.tok.encode_plain(p, max(room, 0)), sp.sep])
        S = min(self.max_len, (max(len(r) for r in rows) + 7) // 8 * 8)
        ids = torch.full((len(rows), S), sp.pad, dtype=torch.int32)
        lens = torch.zeros(len(rows), dtype=torch.int32)
        for i, r in enumerate(rows):
            r = r[:S]
            ids[i, :len(r)] = torch.tensor(r, dtype=torch.int32)
            lens[i] = len(r)
        return ids, lens

    def score(self, query: str, passages: list[str]) -> list[float]:
        if not passages:
            return []
        ids, lens = self.encode_pairs(query, passages)
        ids, lens = ids.to(self.device, non_blocking=True), lens.to(self.device, non_blocking=True)
        logits = self.model.score(ids, lens) if self.device.type == "cuda" else self.model.score_ref(ids, lens)
        return [float(x) for x in logits.float().cpu()]


def rerank_with_cross_encoder(query: str, results: list[RankedResult], reranker: CrossEncoderReranker, *, top_n: int | None = None,
                              max_candidates: int = MAX_RERANK_CANDIDATES) -> list[RankedResult]:
    if not results:
        return results
    head, tail = results[:max_candidates], results[max_candidates:]
    try:
        scores = reranker.score(query, [f"{r.title}. {r.snippet}" for r in head])
    except Exception as exc:  # noqa: BLE001
        logger.warning("cross_encoder_rerank_failed", error=str(exc))
        return results
    order = sorted(range(len(head)), key=lambda i: scores[i], reverse=True)      # stable: ties keep first-stage order
    out = [head[i] for i in order] + tail
    return out[:top_n] if top_n else out
