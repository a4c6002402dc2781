"""Stand-in for the ``trafilatura`` wheel (not installable here: no index access) so that the reference's crawler / services
modules import inside the differential harness.  Extraction is a plain tag strip -- scenarios that go through it compare
bookkeeping (hashes, duplicates, counters), not extraction quality; this package's own extractor is `crawler/parser.py`."""
import re
from types import SimpleNamespace

_DROP = re.compile(r"(?is)<(script|style|noscript|template)\b.*?</\1>")
_TAG = re.compile(r"(?s)<[^>]+>")


def extract(html, url=None, **_kw):
    text = _TAG.sub(" ", _DROP.sub(" ", html or ""))
    text = " ".join(text.split())
    return text or None


def extract_metadata(html, default_url=None, **_kw):
    m = re.search(r"(?is)<title[^>]*>(.*?)</title>", html or "")
    return SimpleNamespace(title=" ".join(m.group(1).split()) if m else None, url=default_url)


def _load_html(html):
    m = re.search(r"(?is)<html\b([^>]*)>", html or "")
    attrs = dict(re.findall(r'([\w:-]+)\s*=\s*["\']([^"\']*)["\']', m.group(1))) if m else {}
    return attrs if m else None


utils = SimpleNamespace(load_html=_load_html)
