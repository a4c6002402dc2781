"""What an LLM client needs on top of ranked hits: context chunks, candidate answers, prompts, entities, a toxicity gate.

Contract (SURVEY §2.1 search/ "RAG helpers"; reference infomesh/search/rag.py:24-413):

* ``format_rag_output``: each hit's snippet is cut into ``chunk_size`` pieces (hit metadata rides only on uncut snippets),
  at most ``max_chunks`` pieces overall; the context window joins ``[Source: title (url)]`` blocks with ``---`` rules.
* ``extract_answers``: snippet sentences (>= 10 characters, each once) that share a word with the query; confidence is
  ``0.8 x query-term coverage + 0.2 x hit score`` capped at 1, kept above 0.2, best ``max_answers`` returned.
* ``build_summary_prompt`` / ``build_cot_rerank_prompt``: fixed prompt frames around numbered results.
* ``extract_entities``: technology names from a fixed vocabulary and capitalised multi-word names (< 30 characters),
  most frequent first.  ``compute_toxicity_score``: flagged words per word x 10, capped at 1.

Implementation: chunking is a generator consumed through ``islice``; sentences, entities and prompts are each one small
pipeline (``Counter`` for the entity tally) so no function carries loop state by hand."""
from __future__ import annotations

import re
from collections import Counter
from dataclasses import asdict, dataclass, field
from itertools import accumulate, islice, takewhile
from typing import Iterator

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
        return asdict(self)

    def cited(self) -> str:
        return f"[Source: {self.title} ({self.url})]\n{self.text}"


@dataclass
class RAGOutput:
    query: str
    chunks: list[RAGChunk]
    total_results: int
    context_window: str = ""

    def to_dict(self) -> dict[str, object]:
        return asdict(self)


_HIT_METADATA = ("bm25_score", "freshness_score", "trust_score", "crawled_at")
_RULE = "\n\n---\n\n"


def _pieces_of(hit: RankedResult, width: int, with_metadata: bool) -> Iterator[RAGChunk]:
    body = hit.snippet or ""
    cuts = [body[at:at + width] for at in range(0, len(body), width)] or [""]
    whole = len(cuts) == 1
    for number, piece in enumerate(cuts):
        details = {name: getattr(hit, name) for name in _HIT_METADATA} if whole and with_metadata else {}
        yield RAGChunk(piece, hit.url, hit.title, hit.combined_score, number, details)


def format_rag_output(query: str, results: list[RankedResult], *, chunk_size: int = 500, max_chunks: int = 10,
                      include_metadata: bool = True) -> RAGOutput:
    stream = (chunk for hit in results for chunk in _pieces_of(hit, chunk_size, include_metadata))
    chunks = list(islice(stream, max_chunks))
    return RAGOutput(query, chunks, len(results), _RULE.join(chunk.cited() for chunk in chunks))


_WORDS = re.compile(r"\w+")
_MARKUP = re.compile(r"</?(?:b|mark|em)>")
_SENTENCE_END = re.compile(r"[.!?]+")


@dataclass(frozen=True)
class ExtractedAnswer:
    answer: str
    source_url: str
    source_title: str
    confidence: float
    context: str = ""


def _vocabulary(text: str) -> frozenset[str]:
    return frozenset(_WORDS.findall(text.lower()))


def extract_answers(query: str, results: list[RankedResult], *, max_answers: int = 3) -> list[ExtractedAnswer]:
    """Sentences sharing terms with the query; confidence = 0.8 * term coverage + 0.2 * result score."""
    wanted = _vocabulary(query)
    per_term = 0.8 / max(len(wanted), 1)
    offered: dict[str, ExtractedAnswer | None] = {}          # sentence -> answer (None: seen but not an answer)
    for hit in results:
        plain = _MARKUP.sub("", hit.snippet or "")           # FTS snippet() highlight markers are not part of the answer
        for sentence in map(str.strip, _SENTENCE_END.split(plain)):
            if len(sentence) < 10 or sentence in offered:
                continue
            shared = len(wanted & _vocabulary(sentence))
            confidence = min(shared * per_term + 0.2 * hit.combined_score, 1.0)
            keep = shared > 0 and confidence > 0.2
            offered[sentence] = ExtractedAnswer(sentence, hit.url, hit.title, round(confidence, 3), plain[:200]) if keep else None
    answers = [a for a in offered.values() if a is not None]
    answers.sort(key=lambda a: a.confidence, reverse=True)
    return answers[:max_answers]


def build_summary_prompt(query: str, results: list[RankedResult], *, max_context: int = 3000) -> str:
    blocks = [f"[{number}] {hit.title}\n{hit.snippet or ''}" for number, hit in enumerate(results, 1)]
    running = accumulate(len(block) for block in blocks)
    fitting = [block for block, _ in takewhile(lambda pair: pair[1] <= max_context, zip(blocks, running))]
    return (f'Based on the following search results for the query "{query}", provide a concise summary that answers the '
            "query.\n\nSearch Results:\n" + "\n\n".join(fitting) + "\n\nSummary:")


@dataclass
class Entity:
    text: str
    entity_type: str        # TECH | NAME
    count: int = 1
    source_urls: list[str] = field(default_factory=list)


_TECH_VOCABULARY = (
    "Python JavaScript TypeScript Rust Go Java C++ Ruby Swift Kotlin React Vue Angular Django Flask FastAPI Node.js Docker "
    "Kubernetes PostgreSQL MySQL Redis MongoDB SQLite AWS Azure GCP Linux macOS Windows GitHub GitLab npm pip cargo CUDA "
    "PyTorch NCCL").split()
_TECH = re.compile(r"\b(?:" + "|".join(map(re.escape, _TECH_VOCABULARY)) + r")\b")
_NAME = re.compile(r"\b([A-Z][a-z]+(?:\s+[A-Z][a-z]+)+)\b")


def extract_entities(text: str, *, source_url: str = "") -> list[Entity]:
    mentions: Counter[tuple[str, str]] = Counter(("TECH", m.group(0)) for m in _TECH.finditer(text))
    mentions.update(("NAME", m.group(1)) for m in _NAME.finditer(text) if len(m.group(1)) < 30)
    origin = [source_url] if source_url else []
    # most_common() is a stable sort on the count, so first-seen order breaks ties (technology names before person names)
    return [Entity(name, kind, times, list(origin)) for (kind, name), times in mentions.most_common()]


_TOXIC = re.compile(r"\b(hate|kill|violence|racist|sexist|porn|gambling|drugs|scam|phishing|malware)\b", re.IGNORECASE)


def compute_toxicity_score(text: str) -> float:
    n_words = len(text.split()) if text else 0
    if not n_words:
        return 0.0
    return min(10.0 * len(_TOXIC.findall(text)) / n_words, 1.0)


def filter_by_toxicity(results: list[RankedResult], *, threshold: float = 0.3) -> list[RankedResult]:
    return [hit for hit in results if compute_toxicity_score(hit.snippet or "") < threshold]


_COT_INSTRUCTIONS = ("For each candidate, think step by step:\n1. What is this result about?\n2. Does it directly answer the query?\n"
                     "3. How relevant is it? (1-10)\n\nThen return a JSON array of objects with \"index\" and \"score\" fields, "
                     "sorted by relevance (highest first).")


def build_cot_rerank_prompt(query: str, results: list[RankedResult], *, max_candidates: int = 10) -> str:
    listing = "\n".join(f"{number}. [{hit.title}] {(hit.snippet or '')[:200]}"
                        for number, hit in enumerate(results[:max_candidates], 1))
    return f'Query: "{query}"\n\nCandidates:\n{listing}\n\n{_COT_INSTRUCTIONS}'
