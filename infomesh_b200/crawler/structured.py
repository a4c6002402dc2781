"""Machine-readable metadata a page carries about itself: JSON-LD blocks, OpenGraph properties, description, keywords.

Contract (SURVEY §2.1 crawler/ "structured data"; reference infomesh/crawler/structured.py): every
``<script type="application/ld+json">`` body that parses contributes its object (or the objects of its top-level list);
``<meta property|name="og:*" content=...>`` fills the OpenGraph map (later duplicates win); the first
``<meta name="description">`` is the description; the first ``<meta name="keywords">`` is split on commas.  Malformed
JSON-LD is skipped, never fatal; an object with nothing in it is falsy.

Implementation: one pass of ``html.parser`` -- attributes are read as a mapping, so their order, quoting and case do not
matter (a pattern over raw markup needs ``property`` before ``content``)."""
from __future__ import annotations

import json
from dataclasses import asdict, dataclass, field
from html.parser import HTMLParser

_LD_TYPE = "application/ld+json"
_OG_PREFIX = "og:"


@dataclass
class StructuredData:
    json_ld: list[dict[str, object]] = field(default_factory=list)
    opengraph: dict[str, str] = field(default_factory=dict)
    meta_description: str = ""
    meta_keywords: list[str] = field(default_factory=list)

    def to_dict(self) -> dict[str, object]:
        return asdict(self)

    def __bool__(self) -> bool:
        return any(asdict(self).values())


class _MetadataReader(HTMLParser):
    def __init__(self):
        super().__init__(convert_charrefs=True)
        self.found = StructuredData()
        self._ld_buffer: list[str] | None = None      # not None while inside a JSON-LD <script>
        self._have_description = self._have_keywords = False

    # ---- <meta>
    def _on_meta(self, attrs: dict[str, str]) -> None:
        key = (attrs.get("property") or attrs.get("name") or "").strip()
        if "content" not in attrs or not key:
            return
        value, folded = attrs["content"], key.lower()
        if folded.startswith(_OG_PREFIX) and len(key) > len(_OG_PREFIX):
            self.found.opengraph[key[len(_OG_PREFIX):]] = value
        elif folded == "description" and "name" in attrs and not self._have_description:
            self.found.meta_description, self._have_description = value.strip(), True
        elif folded == "keywords" and "name" in attrs and not self._have_keywords:
            self.found.meta_keywords = [word for word in (part.strip() for part in value.split(",")) if word]
            self._have_keywords = True

    # ---- <script type="application/ld+json">
    def _close_ld(self) -> None:
        raw, self._ld_buffer = "".join(self._ld_buffer or ()), None
        try:
            doc = json.loads(raw)
        except ValueError:
            return
        # only objects are kept (the field is a list of dicts; a stray scalar inside a top-level list is dropped)
        self.found.json_ld.extend(item for item in (doc if isinstance(doc, list) else [doc]) if isinstance(item, dict))

    def handle_starttag(self, tag, attrs):
        mapping = {name: (value or "") for name, value in attrs}
        if tag == "meta":
            self._on_meta(mapping)
        elif tag == "script" and mapping.get("type", "").strip().lower() == _LD_TYPE:
            self._ld_buffer = []

    handle_startendtag = handle_starttag

    def handle_data(self, data):
        if self._ld_buffer is not None:
            self._ld_buffer.append(data)

    def handle_endtag(self, tag):
        if tag == "script" and self._ld_buffer is not None:
            self._close_ld()


def extract_structured_data(html: str) -> StructuredData:
    reader = _MetadataReader()
    try:
        reader.feed(html)
        reader.close()
    except Exception:  # noqa: BLE001 -- broken markup: keep what was read so far
        pass
    return reader.found
