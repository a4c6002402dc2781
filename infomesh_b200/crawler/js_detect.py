"""Heuristic "does this page need JavaScript?" detector (reference infomesh/crawler/js_detect.py:58-175):
SPA mount point +0.25, framework data blob +0.20, noscript "enable JavaScript" +0.30, text/HTML ratio < 2 % +0.35
(< 5 % +0.15), near-empty body +0.30; js_required when the capped sum >= 0.5."""
from __future__ import annotations

import re
from dataclasses import dataclass


@dataclass(frozen=True)
class JSDetectionResult:
    js_required: bool
    confidence: float
    signals: list[str]


_SPA_ROOT = re.compile(r'<div\s+id=["\'](?:root|app|__next|__nuxt|__vue)["\']', re.I)
_FRAMEWORK = re.compile(r"__NEXT_DATA__|__NUXT__|window\.__INITIAL_STATE__|window\.webpackJsonp|window\.__remixContext",
                        re.I)
_NOSCRIPT = re.compile(r"<noscript[^>]*>(.*?)</noscript>", re.I | re.S)
_NEEDS_JS = re.compile(r"enable\s+javascript|javascript\s+(?:is\s+)?required|(?:need|requires?)\s+javascript|"
                       r"must\s+enable\s+javascript|activate\s+javascript|turn\s+on\s+javascript", re.I)
_BODY = re.compile(r"<body[^>]*>(.*?)</body>", re.I | re.S)
_TAG = re.compile(r"<[^>]+>")
_SCRIPT = re.compile(r"<script[^>]*>.*?</script>", re.I | re.S)
_STYLE = re.compile(r"<style[^>]*>.*?</style>", re.I | re.S)
_WS = re.compile(r"\s+")


def _visible_text(fragment: str) -> str:
    return _WS.sub(" ", _TAG.sub("", _STYLE.sub("", _SCRIPT.sub("", fragment)))).strip()


def _text_to_html_ratio(html: str) -> float:
    return len(_visible_text(html)) / len(html) if html else 0.0


def _body_text_length(html: str) -> int:
    m = _BODY.search(html)
    return len(_visible_text(m.group(1) if m else html))


def detect_js_requirement(html: str) -> JSDetectionResult:
    signals: list[str] = []
    score = 0.0
    if _SPA_ROOT.search(html):
        signals.append("SPA root element detected (div#root, #app, #__next, …)")
        score += 0.25
    if _FRAMEWORK.search(html):
        signals.append("Framework data blob detected (__NEXT_DATA__, etc.)")
        score += 0.20
    if any(_NEEDS_JS.search(m.group(1)) for m in _NOSCRIPT.finditer(html)):
        signals.append("Noscript fallback says JavaScript is required")
        score += 0.30
    ratio = _text_to_html_ratio(html)
    if ratio < 0.02:
        signals.append(f"Very low text-to-HTML ratio ({ratio:.1%})")
        score += 0.35
    elif ratio < 0.05:
        signals.append(f"Low text-to-HTML ratio ({ratio:.1%})")
        score += 0.15
    body_len = _body_text_length(html)
    if body_len < 50:
        signals.append(f"Near-empty body ({body_len} chars of text)")
        score += 0.30
    conf = min(score, 1.0)
    return JSDetectionResult(conf >= 0.5, round(conf, 2), signals)
