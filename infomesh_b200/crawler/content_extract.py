"""Code-block and table extraction from raw HTML (reference infomesh/crawler/content_extract.py:29-196)."""
from __future__ import annotations

import html as _html
import re
from dataclasses import dataclass, field

_CODE = re.compile(r'<pre[^>]*>\s*<code(?:\s+class=["\'](?:language-|lang-)?([\w+#-]+)["\'])?[^>]*>(.*?)</code>\s*</pre>',
                   re.S | re.I)
_TABLE = re.compile(r"<table[^>]*>(.*?)</table>", re.S | re.I)
_TR = re.compile(r"<tr[^>]*>(.*?)</tr>", re.S | re.I)
_CELL = re.compile(r"<(?:th|td)[^>]*>(.*?)</(?:th|td)>", re.S | re.I)
_CAPTION = re.compile(r"<caption[^>]*>(.*?)</caption>", re.S | re.I)
_TAG = re.compile(r"<[^>]+>")


def _clean_html(text: str) -> str:
    return _html.unescape(_TAG.sub("", text)).replace("\xa0", " ")


@dataclass(frozen=True)
class CodeBlock:
    code: str
    language: str = ""
    line_count: int = 0


def extract_code_blocks(html: str, *, min_lines: int = 2, max_blocks: int = 50) -> list[CodeBlock]:
    out: list[CodeBlock] = []
    for m in _CODE.finditer(html):
        if len(out) >= max_blocks:
            break
        code = _clean_html(m.group(2))
        n = code.count("\n") + 1
        if n >= min_lines:
            out.append(CodeBlock(code, (m.group(1) or "").lower(), n))
    return out


@dataclass
class ExtractedTable:
    headers: list[str] = field(default_factory=list)
    rows: list[list[str]] = field(default_factory=list)
    caption: str = ""

    def to_csv(self) -> str:
        q = lambda cells: ",".join('"' + c.replace('"', '""') + '"' for c in cells)  # noqa: E731
        return "\n".join(([q(self.headers)] if self.headers else []) + [q(r) for r in self.rows])

    def to_dict_list(self) -> list[dict[str, str]]:
        if not self.headers:
            return []
        return [{h: (row[i] if i < len(row) else "") for i, h in enumerate(self.headers)} for row in self.rows]


def extract_tables(html: str, *, max_tables: int = 20, min_rows: int = 1) -> list[ExtractedTable]:
    out: list[ExtractedTable] = []
    for tm in _TABLE.finditer(html):
        if len(out) >= max_tables:
            break
        body = tm.group(1)
        cap = _CAPTION.search(body)
        headers: list[str] = []
        rows: list[list[str]] = []
        for tr in _TR.finditer(body):
            cells = [_clean_html(c.group(1)).strip() for c in _CELL.finditer(tr.group(1))]
            if not cells:
                continue
            if not headers and "<th" in tr.group(1).lower():
                headers = cells
            else:
                rows.append(cells)
        if len(rows) >= min_rows:
            out.append(ExtractedTable(headers, rows, _clean_html(cap.group(1)).strip() if cap else ""))
    return out
