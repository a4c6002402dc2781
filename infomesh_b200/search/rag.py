"""RAG-side helpers on top of ranked results: chunked context windows, sentence-level answer extraction, multi-result
summary prompts, pattern entity extraction, keyword toxicity filter, chain-of-thought rerank prompts
(reference infomesh/search/rag.py:24-413)."""
from __future__ import annotations

import re
from dataclasses import dataclass, field

from infomesh_b200.index.ranking import RankedResult


@dataclass
class RAGChunk:
    text: str
    url: str
    title: str
    score: float
    chunk_index: int = 0
    metadata: dict[str, object] = field(default_factory=dict)

    def to_dict(self) -> dict[str, object]:
        return {"text": self.text, "url": self.url, "title": self.title, "score": self.score, "chunk_index": self.chunk_index,
                "metadata": self.metadata}


@dataclass
class RAGOutput:
    query: str
    chunks: list[RAGChunk]
    total_results: int
    context_window: str = ""

    def to_dict(self) -> dict[str, object]:
        return {"query": self.query, "chunks": [c.to_dict() for c in self.chunks], "total_results": self.total_results,
                "context_window": self.context_window}


def format_rag_output(query: str, results: list[RankedResult], *, chunk_size: int = 500, max_chunks: int = 10,
                      include_metadata: bool = True) -> RAGOutput:
    chunks: list[RAGChunk] = []
    for r in results:
        text = r.snippet or ""
        meta = ({"bm25_score": r.bm25_score, "freshness_score": r.freshness_score, "trust_score": r.trust_score,
                 "crawled_at": r.crawled_at} if include_metadata else {})
        pieces = [text] if len(text) <= chunk_size else [text[i:i + chunk_size] for i in range(0, len(text), chunk_size)]
        for i, piece in enumerate(pieces):
            chunks.append(RAGChunk(piece, r.url, r.title, r.combined_score, i, meta if len(pieces) == 1 else {}))
        if len(chunks) >= max_chunks:
            break
    chunks = chunks[:max_chunks]
    window = "\n\n---\n\n".join(f"[Source: {c.title} ({c.url})]\n{c.text}" for c in chunks)
    return RAGOutput(query, chunks, len(results), window)


_WORDS = re.compile(r"\w+")
_MARKUP = re.compile(r"</?(?:b|mark|em)>")


@dataclass(frozen=True)
class ExtractedAnswer:
    answer: str
    source_url: str
    source_title: str
    confidence: float
    context: str = ""


def extract_answers(query: str, results: list[RankedResult], *, max_answers: int = 3) -> list[ExtractedAnswer]:
    """Sentences sharing terms with the query; confidence = 0.8 * term coverage + 0.2 * result score."""
    terms = set(_WORDS.findall(query.lower()))
    found: list[ExtractedAnswer] = []
    seen: set[str] = set()
    for r in results:
        text = _MARKUP.sub("", r.snippet or "")          # FTS snippet() highlight markers are not part of the answer
        for sent in (s.strip() for s in re.split(r"[.!?]+", text)):
            if len(sent) < 10 or sent in seen:
                continue
            seen.add(sent)
            overlap = len(terms & set(_WORDS.findall(sent.lower())))
            if not overlap:
                continue
            conf = min(overlap / max(len(terms), 1) * 0.8 + r.combined_score * 0.2, 1.0)
            if conf > 0.2:
                found.append(ExtractedAnswer(sent, r.url, r.title, round(conf, 3), text[:200]))
    return sorted(found, key=lambda a: a.confidence, reverse=True)[:max_answers]


def build_summary_prompt(query: str, results: list[RankedResult], *, max_context: int = 3000) -> str:
    parts, used = [], 0
    for i, r in enumerate(results, 1):
        entry = f"[{i}] {r.title}\n{r.snippet or ''}"
        if used + len(entry) > max_context:
            break
        parts.append(entry)
        used += len(entry)
    ctx = "\n\n".join(parts)
    return (f'Based on the following search results for the query "{query}", provide a concise summary that answers the '
            f"query.\n\nSearch Results:\n{ctx}\n\nSummary:")


@dataclass
class Entity:
    text: str
    entity_type: str        # TECH | NAME
    count: int = 1
    source_urls: list[str] = field(default_factory=list)


_TECH = re.compile(r"\b(?:Python|JavaScript|TypeScript|Rust|Go|Java|C\+\+|Ruby|Swift|Kotlin|React|Vue|Angular|Django|Flask|FastAPI|"
                   r"Node\.js|Docker|Kubernetes|PostgreSQL|MySQL|Redis|MongoDB|SQLite|AWS|Azure|GCP|Linux|macOS|Windows|GitHub|"
                   r"GitLab|npm|pip|cargo|CUDA|PyTorch|NCCL)\b")
_NAME = re.compile(r"\b([A-Z][a-z]+(?:\s+[A-Z][a-z]+)+)\b")


def extract_entities(text: str, *, source_url: str = "") -> list[Entity]:
    bag: dict[tuple[str, str], Entity] = {}

    def note(kind: str, name: str) -> None:
        e = bag.get((kind, name))
        if e:
            e.count += 1
        else:
            bag[(kind, name)] = Entity(name, kind, 1, [source_url] if source_url else [])

    for m in _TECH.finditer(text):
        note("TECH", m.group(0))
    for m in _NAME.finditer(text):
        if len(m.group(1)) < 30:
            note("NAME", m.group(1))
    return sorted(bag.values(), key=lambda e: e.count, reverse=True)


_TOXIC = re.compile(r"\b(hate|kill|violence|racist|sexist|porn|gambling|drugs|scam|phishing|malware)\b", re.IGNORECASE)


def compute_toxicity_score(text: str) -> float:
    words = len(text.split()) if text else 0
    return min(len(_TOXIC.findall(text)) / words * 10, 1.0) if words else 0.0


def filter_by_toxicity(results: list[RankedResult], *, threshold: float = 0.3) -> list[RankedResult]:
    return [r for r in results if compute_toxicity_score(r.snippet or "") < threshold]


def build_cot_rerank_prompt(query: str, results: list[RankedResult], *, max_candidates: int = 10) -> str:
    cands = "\n".join(f"{i}. [{r.title}] {(r.snippet or '')[:200]}" for i, r in enumerate(results[:max_candidates], 1))
    return (f'Query: "{query}"\n\nCandidates:\n{cands}\n\nFor each candidate, think step by step:\n1. What is this result about?\n'
            "2. Does it directly answer the query?\n3. How relevant is it? (1-10)\n\nThen return a JSON array of objects with "
            '"index" and "score" fields, sorted by relevance (highest first).')
