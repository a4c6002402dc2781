"""Structured metadata embedded in pages: JSON-LD, OpenGraph, meta description / keywords
(reference infomesh/crawler/structured.py:14-92)."""
from __future__ import annotations

import json
import re
from dataclasses import dataclass, field

_JSON_LD = re.compile(r'<script[^>]*type=["\']application/ld\+json["\'][^>]*>(.*?)</script>', re.S | re.I)
_OG = re.compile(r'<meta\s+(?:property|name)=["\']og:([^"\']+)["\']\s+content=["\']([^"\']*)["\']', re.I)
_DESC = re.compile(r'<meta\s+name=["\']description["\']\s+content=["\']([^"\']*)["\']', re.I)
_KEYWORDS = re.compile(r'<meta\s+name=["\']keywords["\']\s+content=["\']([^"\']*)["\']', re.I)


@dataclass
class StructuredData:
    json_ld: list[dict[str, object]] = field(default_factory=list)
    opengraph: dict[str, str] = field(default_factory=dict)
    meta_description: str = ""
    meta_keywords: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, object]:
        return {"json_ld": self.json_ld, "opengraph": self.opengraph, "meta_description": self.meta_description,
                "meta_keywords": self.meta_keywords}

    def __bool__(self) -> bool:
        return bool(self.json_ld or self.opengraph or self.meta_description or self.meta_keywords)


def extract_structured_data(html: str) -> StructuredData:
    out = StructuredData()
    for m in _JSON_LD.finditer(html):
        try:
            data = json.loads(m.group(1))
        except ValueError:
            continue
        if isinstance(data, list):
            out.json_ld.extend(d for d in data if isinstance(d, dict))
        elif isinstance(data, dict):
            out.json_ld.append(data)
    for m in _OG.finditer(html):
        out.opengraph[m.group(1)] = m.group(2)
    d = _DESC.search(html)
    if d:
        out.meta_description = d.group(1).strip()
    k = _KEYWORDS.search(html)
    if k:
        out.meta_keywords = [w.strip() for w in k.group(1).split(",") if w.strip()]
    return out
