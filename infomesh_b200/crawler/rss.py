"""Dependency-free RSS 2.0 / Atom parsing and feed discovery (reference infomesh/crawler/rss.py:13-197)."""
from __future__ import annotations

import html as _html
import re
from dataclasses import dataclass, field
from urllib.parse import urljoin


@dataclass(frozen=True)
class FeedItem:
    title: str
    url: str
    summary: str = ""
    published: str = ""
    author: str = ""


@dataclass
class FeedResult:
    title: str
    url: str
    items: list[FeedItem] = field(default_factory=list)
    feed_type: str = "unknown"  # rss | atom


_FEED_LINK = re.compile(r'<link[^>]*type=["\']application/(?:rss|atom)\+xml["\'][^>]*>', re.I | re.S)
_HREF = re.compile(r'href=["\']([^"\']+)["\']', re.I)
_CDATA = re.compile(r"<!\[CDATA\[(.*?)\]\]>", re.S)
_TAGS = re.compile(r"<[^>]+>")


def discover_feeds(html: str, base_url: str) -> list[str]:
    out: list[str] = []
    for m in _FEED_LINK.finditer(html):
        h = _HREF.search(m.group(0))
        if h:
            url = urljoin(base_url, h.group(1))
            if url not in out:
                out.append(url)
    return out


def _strip_cdata(text: str) -> str:
    return _CDATA.sub(lambda m: m.group(1), text).strip()


def _tag(block: str, name: str) -> str:
    m = re.search(rf"<{name}(?:\s[^>]*)?>(.*?)</{name}>", block, re.I | re.S)
    return _html.unescape(_strip_cdata(m.group(1))) if m else ""


def _plain(text: str, limit: int = 500) -> str:
    return re.sub(r"\s+", " ", _TAGS.sub("", text)).strip()[:limit]


def _parse_item(block: str, feed_type: str) -> FeedItem:
    title = _plain(_tag(block, "title"), 300)
    if feed_type == "atom":
        m = (re.search(r'<link[^>]*rel=["\']alternate["\'][^>]*href=["\']([^"\']+)["\']', block, re.I)
             or re.search(r'<link[^>]*href=["\']([^"\']+)["\']', block, re.I))
        url = _html.unescape(m.group(1)) if m else ""
        summary = _tag(block, "summary") or _tag(block, "content")
        published = _tag(block, "published") or _tag(block, "updated")
        author = _plain(_tag(_tag_raw(block, "author"), "name") or "", 100)
    else:
        url = _tag(block, "link") or _tag(block, "guid")
        summary = _tag(block, "description")
        published = _tag(block, "pubDate") or _tag(block, "dc:date")
        author = _plain(_tag(block, "author") or _tag(block, "dc:creator"), 100)
    return FeedItem(title, url.strip(), _plain(summary), published.strip(), author)


def _tag_raw(block: str, name: str) -> str:
    m = re.search(rf"<{name}(?:\s[^>]*)?>(.*?)</{name}>", block, re.I | re.S)
    return m.group(0) if m else ""


def parse_feed_xml(xml_text: str, feed_url: str) -> FeedResult:
    is_atom = bool(re.search(r"<feed[\s>]", xml_text, re.I)) and not re.search(r"<rss[\s>]", xml_text, re.I)
    kind = "atom" if is_atom else ("rss" if re.search(r"<(rss|rdf:RDF|channel)[\s>]", xml_text, re.I) else "unknown")
    item_tag = "entry" if is_atom else "item"
    head = re.split(rf"<{item_tag}[\s>]", xml_text, maxsplit=1, flags=re.I)[0]
    res = FeedResult(title=_plain(_tag(head, "title"), 300), url=feed_url, feed_type=kind)
    for m in re.finditer(rf"<{item_tag}(?:\s[^>]*)?>(.*?)</{item_tag}>", xml_text, re.I | re.S):
        it = _parse_item(m.group(1), "atom" if is_atom else "rss")
        if it.url:
            res.items.append(FeedItem(it.title, urljoin(feed_url, it.url), it.summary, it.published, it.author))
    return res
