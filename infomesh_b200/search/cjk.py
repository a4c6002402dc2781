"""CJK / Thai aware query handling: script detection, character n-grams, tokenizer recommendation,
lightweight Chinese / Korean segmentation (reference infomesh/search/cjk.py:54-261)."""
from __future__ import annotations

import re

_CJK = re.compile("[\\u4e00-\\u9fff\\u3400-\\u4dbf\\uf900-\\ufaff]")          # ideographs (+ ext. A, compatibility)
_HANGUL = re.compile("[\\uac00-\\ud7af\\u1100-\\u11ff\\u3130-\\u318f]")       # syllables, jamo, compatibility jamo
_KANA = re.compile("[\\u3040-\\u309f\\u30a0-\\u30ff]")
_THAI = re.compile("[\\u0e00-\\u0e7f]")
_ARABIC = re.compile("[\\u0600-\\u06ff]")
_DEVANAGARI = re.compile("[\\u0900-\\u097f]")
_NGRAM_SCRIPTS = (_CJK, _HANGUL, _KANA, _THAI)


def _is_ngram_char(ch: str) -> bool:
    return any(p.match(ch) for p in _NGRAM_SCRIPTS)


def is_cjk_text(text: str, threshold: float = 0.3) -> bool:
    """True when CJK ideographs + Hangul + Kana + Thai make up at least ``threshold`` of the non-space chars."""
    if not text:
        return False
    total = sum(1 for ch in text if not ch.isspace())
    if total == 0:
        return False
    hits = sum(len(p.findall(text)) for p in _NGRAM_SCRIPTS)
    return hits / total >= threshold


def _ngrams(chars: list[str], n: int) -> list[str]:
    if len(chars) < n:
        return ["".join(chars)]
    return ["".join(chars[i:i + n]) for i in range(len(chars) - n + 1)]


def _char_ngrams(text: str, n: int) -> list[str]:
    """n-grams over runs of CJK-like characters; ASCII alphanumeric runs are kept whole."""
    out: list[str] = []
    run: list[str] = []
    kind = ""  # "c" (cjk) | "l" (latin)

    def flush() -> None:
        nonlocal run, kind
        if run:
            out.extend(_ngrams(run, n) if kind == "c" else ["".join(run)])
        run, kind = [], ""

    for ch in text:
        k = "c" if _is_ngram_char(ch) else ("l" if ch.isascii() and ch.isalnum() else "")
        if k != kind:
            flush()
        if k:
            kind = k
            run.append(ch)
    flush()
    return out


def cjk_bigrams(text: str) -> list[str]:
    return _char_ngrams(text, 2)


def cjk_trigrams(text: str) -> list[str]:
    return _char_ngrams(text, 3)


def recommend_tokenizer(sample_text: str) -> str:
    """``trigram`` for CJK-heavy corpora, ``unicode61`` otherwise."""
    return "trigram" if is_cjk_text(sample_text, threshold=0.2) else "unicode61"


def tokenize_query_cjk(query: str) -> str:
    """Queries with >= 20 % CJK-like characters are rewritten as space-joined bigrams."""
    if not is_cjk_text(query, threshold=0.2):
        return query
    toks = cjk_bigrams(query)
    return " ".join(toks) if toks else query


def segment_chinese(text: str) -> list[str]:
    try:
        import jieba  # type: ignore

        return list(jieba.cut(text))
    except ImportError:
        return cjk_bigrams(text)


def segment_korean(text: str) -> list[str]:
    """Hangul runs of <= 4 syllables are kept whole, longer runs become bigrams; Latin runs stay intact."""
    out: list[str] = []
    hangul: list[str] = []
    latin: list[str] = []

    def flush_h() -> None:
        if hangul:
            word = "".join(hangul)
            out.extend([word] if len(word) <= 4 else _ngrams(list(word), 2))
            hangul.clear()

    def flush_l() -> None:
        if latin:
            out.append("".join(latin))
            latin.clear()

    for ch in text:
        if _HANGUL.match(ch):
            flush_l()
            hangul.append(ch)
        elif ch.isalnum():
            flush_h()
            latin.append(ch)
        else:
            flush_h()
            flush_l()
    flush_h()
    flush_l()
    return out


def detect_script(text: str) -> str:
    """Dominant script label: cjk | hangul | kana | thai | arabic | devanagari | latin."""
    counts = {"cjk": len(_CJK.findall(text)), "hangul": len(_HANGUL.findall(text)), "kana": len(_KANA.findall(text)),
              "thai": len(_THAI.findall(text)), "arabic": len(_ARABIC.findall(text)),
              "devanagari": len(_DEVANAGARI.findall(text))}
    best = max(counts, key=lambda k: counts[k])
    return best if counts[best] > 0 else "latin"
