"""HTML -> (title, main text, language, hashes, image alts), canonical URL and outbound links.

The reference delegates extraction to trafilatura (infomesh/crawler/parser.py:36-126); trafilatura / lxml are not
available offline, so the default extractor here is a stdlib ``html.parser`` readability pass (drop
script/style/nav/aside/footer/form chrome, keep block text and tables, prefer <article>/<main>), and trafilatura is
used instead whenever it can be imported.  Contract kept: >= 50 chars of text or ``None``; title from metadata /
<title>; language from ``<html lang>`` then ``detect_language`` (confidence > 0.3); SHA-256 hashes.
"""
from __future__ import annotations

import html as _html
import re
from dataclasses import dataclass
from html.parser import HTMLParser
from urllib.parse import urljoin, urlparse

from infomesh_b200.hashing import content_hash
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

MIN_TEXT_CHARS = 50


@dataclass(frozen=True)
class ParsedPage:
    url: str
    title: str
    text: str
    language: str | None
    raw_html_hash: str
    text_hash: str
    image_alt_texts: list[str] = ()  # type: ignore[assignment]


_DROP = frozenset({"script", "style", "noscript", "template", "svg", "canvas", "iframe", "object", "embed", "head",
                   "nav", "aside", "footer", "form", "button", "select", "option", "dialog", "menu"})
_BLOCK = frozenset({"p", "div", "section", "article", "main", "li", "ul", "ol", "h1", "h2", "h3", "h4", "h5", "h6",
                    "blockquote", "pre", "table", "tr", "br", "hr", "dd", "dt", "figcaption", "header", "details",
                    "summary", "address"})
_VOID = frozenset({"br", "hr", "img", "input", "meta", "link", "area", "base", "col", "embed", "source", "track", "wbr"})
_MAIN = frozenset({"article", "main"})
_IMG_ALT_RE = re.compile(r'<img\s[^>]*alt=["\']([^"\']+)["\']', re.I)
_TITLE_RE = re.compile(r"<title[^>]*>(.*?)</title>", re.I | re.S)
_LANG_RE = re.compile(r"<html[^>]*\s(?:xml:)?lang=[\"']?([a-zA-Z-]{2,10})", re.I)
_OG_TITLE_RE = re.compile(r'<meta\s[^>]*property=["\']og:title["\'][^>]*content=["\']([^"\']+)["\']', re.I)
_WS = re.compile(r"[ \t\r\f\v]+")


class _TextExtractor(HTMLParser):
    def __init__(self):
        super().__init__(convert_charrefs=True)
        self.drop_depth = 0
        self.main_depth = 0
        self.all: list[str] = []
        self.main: list[str] = []
        self._stack: list[str] = []

    def _emit(self, s: str) -> None:
        self.all.append(s)
        if self.main_depth:
            self.main.append(s)

    def handle_starttag(self, tag, attrs):
        if tag in _VOID:
            if tag in ("br", "hr") and not self.drop_depth:
                self._emit("\n")
            return
        self._stack.append(tag)
        if tag in _DROP:
            self.drop_depth += 1
        elif tag in _MAIN:
            self.main_depth += 1
        if not self.drop_depth:
            if tag in _BLOCK:
                self._emit("\n")
            elif tag in ("td", "th"):
                self._emit(" | ")

    def handle_endtag(self, tag):
        if tag in _VOID or tag not in self._stack:
            return
        while self._stack:
            t = self._stack.pop()
            if t in _DROP:
                self.drop_depth = max(0, self.drop_depth - 1)
            elif t in _MAIN:
                self.main_depth = max(0, self.main_depth - 1)
            if t == tag:
                break
        if not self.drop_depth and tag in _BLOCK:
            self._emit("\n")

    def handle_data(self, data):
        if not self.drop_depth and data:
            self._emit(data)


def _clean(parts: list[str]) -> str:
    lines = []
    for ln in "".join(parts).split("\n"):
        ln = _WS.sub(" ", ln).strip(" |")
        if ln:
            lines.append(ln)
    out: list[str] = []
    for ln in lines:   # collapse runs of identical boilerplate lines
        if not out or out[-1] != ln:
            out.append(ln)
    return "\n\n".join(out)


def extract_main_text(html: str) -> str:
    """Stdlib readability pass: prefer <article>/<main> when it carries most of the text."""
    p = _TextExtractor()
    try:
        p.feed(html)
        p.close()
    except Exception:  # noqa: BLE001 — malformed markup: keep whatever was parsed
        pass
    whole, main = _clean(p.all), _clean(p.main)
    return main if len(main) >= max(200, 0.3 * len(whole)) else whole


def _extract_title(html: str) -> str:
    m = _OG_TITLE_RE.search(html) or _TITLE_RE.search(html)
    return _WS.sub(" ", _html.unescape(m.group(1))).strip() if m else ""


def _extract_image_alts(html: str) -> list[str]:
    seen: set[str] = set()
    out: list[str] = []
    for m in _IMG_ALT_RE.finditer(html):
        alt = m.group(1).strip()
        if len(alt) > 3 and alt not in seen:
            seen.add(alt)
            out.append(alt)
    return out


def extract_content(html: str, url: str, *, raw_hash: str = "") -> ParsedPage | None:
    try:
        text, title = "", ""
        try:  # optional high-quality extractor
            import trafilatura  # type: ignore

            text = trafilatura.extract(html, url=url, include_links=False, include_images=False,
                                       include_tables=True, output_format="txt", favor_recall=True) or ""
            meta = trafilatura.extract_metadata(html, default_url=url)
            title = (meta.title if meta and meta.title else "") or ""
        except ImportError:
            text = extract_main_text(html)
        text = text.strip()
        if len(text) < MIN_TEXT_CHARS:
            logger.debug("parse_empty", url=url)
            return None
        title = title or _extract_title(html)
        language = None
        m = _LANG_RE.search(html[:4096])
        if m:
            language = m.group(1)[:2].lower()
        if not language:
            from infomesh_b200.crawler.lang_detect import detect_language

            det = detect_language(text[:2000])
            if det.confidence > 0.3:
                language = det.language
        return ParsedPage(url=url, title=title, text=text, language=language,
                          raw_html_hash=raw_hash or content_hash(html), text_hash=content_hash(text),
                          image_alt_texts=_extract_image_alts(html))
    except Exception as exc:  # noqa: BLE001
        logger.error("parse_error", url=url, error=str(exc))
        return None


_HREF_RE = re.compile(r"<a\s[^>]*href=(?:[\"']([^\"'#][^\"']*)[\"']|([^\s>\"'#][^\s>\"']*))", re.I)
_CANONICAL_RES = (re.compile(r'<link\s[^>]*rel=["\']canonical["\'][^>]*href=["\']([^"\']+)["\']', re.I),
                  re.compile(r'<link\s[^>]*href=["\']([^"\']+)["\'][^>]*rel=["\']canonical["\']', re.I))
_SKIP_EXTENSIONS = frozenset({".pdf", ".jpg", ".jpeg", ".png", ".gif", ".svg", ".webp", ".mp3", ".mp4", ".avi",
                              ".mov", ".zip", ".tar", ".gz", ".exe", ".dmg", ".iso", ".css", ".js", ".woff", ".woff2"})
_NON_HTTP = ("mailto:", "javascript:", "tel:", "data:")


def extract_canonical(html: str, page_url: str) -> str | None:
    for rx in _CANONICAL_RES:
        m = rx.search(html)
        if m and m.group(1).strip():
            absolute = urljoin(page_url, m.group(1).strip())
            return absolute if urlparse(absolute).scheme in ("http", "https") else None
    return None


def extract_links(html: str, base_url: str) -> list[str]:
    """Absolute http(s) links in document order, fragments stripped, binary / asset extensions skipped."""
    seen: set[str] = set()
    out: list[str] = []
    for m in _HREF_RE.finditer(html):
        href = _html.unescape((m.group(1) or m.group(2) or "").strip())
        if not href or href.lower().startswith(_NON_HTTP):
            continue
        try:
            absolute = urljoin(base_url, href)
            p = urlparse(absolute)
        except ValueError:
            continue
        if p.scheme not in ("http", "https"):
            continue
        path = p.path.lower()
        if any(path.endswith(ext) for ext in _SKIP_EXTENSIONS):
            continue
        clean = absolute.split("#", 1)[0]
        if clean and clean not in seen:
            seen.add(clean)
            out.append(clean)
    return out
