"""Passage splitting, scoring, highlighting, title/URL match helpers and intent classification.

Semantics follow reference infomesh/search/passage.py:39-374 (paragraph -> sentence -> word fallback with
500/40 char bounds; score = coverage + 0.1 * density; ``<b>`` highlighting; title / URL-path overlap in [0, 1];
three-way intent).  The GPU twin of ``select_best_passage`` is ``ops.bm25.passage_score`` (K11).
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from urllib.parse import urlparse

_PARA = re.compile(r"\n\s*\n")
_SENT = re.compile(r"(?<=[.!?])\s+(?=[A-Zㄱ-ㆎ一-鿿])")
_TOKEN = re.compile(r"[a-zA-Z0-9ㄱ-ㆎ가-힣一-鿿]+")
_PATH_WORD = re.compile(r"[a-z0-9]+")


@dataclass(frozen=True)
class ScoredPassage:
    text: str
    score: float
    start: int
    end: int


def _tokenize(text: str) -> list[str]:
    """Lower-cased alphanumeric / Hangul / CJK runs."""
    return _TOKEN.findall(text.lower())


tokenize = _tokenize


def _pack(pieces: list[str], max_length: int) -> list[str]:
    """Greedily join pieces with single spaces without exceeding ``max_length`` (a piece alone may exceed)."""
    out: list[str] = []
    cur: list[str] = []
    size = 0
    for piece in pieces:
        if cur and size + len(piece) > max_length:
            out.append(" ".join(cur))
            cur, size = [], 0
        cur.append(piece)
        size += len(piece) + 1
    if cur:
        out.append(" ".join(cur))
    return out


def _split_long(chunk: str, max_length: int, min_length: int, out: list[str]) -> None:
    sentences = [s.strip() for s in _SENT.split(chunk)]
    sentences = [s for s in sentences if s]
    if len(sentences) <= 1:
        out.extend(_pack(chunk.split(), max_length))
        return
    packed = _pack(sentences, max_length)
    # a short trailing remainder is folded into its predecessor
    if len(packed) >= 1:
        last = packed[-1]
        if len(last) < min_length and (out or len(packed) > 1):
            packed.pop()
            if packed:
                packed[-1] = packed[-1] + " " + last
            elif out:
                out[-1] = out[-1] + " " + last
    out.extend(packed)


def split_passages(text: str, *, max_length: int = 500, min_length: int = 40) -> list[str]:
    if not text or not text.strip():
        return []
    passages: list[str] = []
    for raw in _PARA.split(text):
        chunk = raw.strip()
        if len(chunk) < min_length:
            if passages:
                passages[-1] = passages[-1] + " " + chunk
            continue
        if len(chunk) <= max_length:
            passages.append(chunk)
        else:
            _split_long(chunk, max_length, min_length, passages)
    return passages


def score_passage(passage: str, query_tokens: list[str]) -> float:
    if not passage or not query_tokens:
        return 0.0
    toks = _tokenize(passage)
    if not toks:
        return 0.0
    wanted = set(query_tokens)
    coverage = len(wanted.intersection(toks)) / len(wanted)
    density = sum(1 for t in toks if t in wanted) / len(toks)
    return coverage + 0.1 * density


def rank_passages(text: str, query: str, *, max_length: int = 500) -> list[ScoredPassage]:
    """All passages with scores and character offsets, best first."""
    q = _tokenize(query)
    res: list[ScoredPassage] = []
    cursor = 0
    for p in split_passages(text, max_length=max_length):
        head = p[:20]
        at = text.find(head, cursor)
        if at < 0:
            at = cursor
        res.append(ScoredPassage(p, score_passage(p, q), at, at + len(p)))
        cursor = at
    res.sort(key=lambda s: s.score, reverse=True)
    return res


def select_best_passage(text: str, query: str, *, max_length: int = 300, fallback_length: int = 200) -> str:
    if not text:
        return ""
    if not query:
        return text[:fallback_length]
    q = _tokenize(query)
    if not q:
        return text[:fallback_length]
    best, best_score = "", -1.0
    for p in split_passages(text, max_length=max_length):
        s = score_passage(p, q)
        if s > best_score:
            best, best_score = p, s
    if best_score <= 0:
        return text[:fallback_length]
    return best[:max_length]


def highlight_terms(text: str, query_tokens: list[str]) -> str:
    if not text or not query_tokens:
        return text
    alts = [re.escape(t) for t in sorted(set(query_tokens), key=len, reverse=True) if t]
    if not alts:
        return text
    return re.sub(rf"\b({'|'.join(alts)})\b", r"<b>\1</b>", text, flags=re.IGNORECASE)


def title_match_score(title: str, query_tokens: list[str]) -> float:
    wanted = set(query_tokens or ())
    if not title or not wanted:
        return 0.0
    return len(wanted.intersection(_tokenize(title))) / len(wanted)


def url_path_score(url: str, query_tokens: list[str]) -> float:
    wanted = set(query_tokens or ())
    if not url or not wanted:
        return 0.0
    try:
        path = urlparse(url).path.lower()
    except ValueError:
        return 0.0
    if path in ("", "/"):
        return 0.0
    words = _PATH_WORD.findall(path)
    if not words:
        return 0.0
    joined = " ".join(words)
    return sum(1 for t in wanted if t in joined) / len(wanted)


class QueryIntent:
    INFORMATIONAL = "informational"
    NAVIGATIONAL = "navigational"
    TRANSACTIONAL = "transactional"


_NAV = (re.compile(r"\b(login|signin|sign\s+in|homepage|official)\b", re.I),
        re.compile(r"\b(go\s+to|open|visit|navigate)\b", re.I),
        re.compile(r"^[a-zA-Z0-9.-]+\.(com|org|net|io|dev|edu|gov)$"))
_TRANS = (re.compile(r"\b(download|install|buy|purchase|subscribe|pricing)\b", re.I),
          re.compile(r"\b(free|trial|demo|signup|register)\b", re.I))


def classify_intent(query: str) -> str:
    if query:
        if any(p.search(query) for p in _NAV):
            return QueryIntent.NAVIGATIONAL
        if any(p.search(query) for p in _TRANS):
            return QueryIntent.TRANSACTIONAL
    return QueryIntent.INFORMATIONAL
