"""PDF text extraction through PyMuPDF when it is installed (reference infomesh/crawler/pdf.py:16-71)."""
from __future__ import annotations

from dataclasses import dataclass

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)


@dataclass(frozen=True)
class PDFContent:
    text: str
    title: str
    page_count: int
    metadata: dict[str, str]


def is_pdf_url(url: str) -> bool:
    low = url.lower().rstrip("/")
    return low.endswith(".pdf") or "application/pdf" in low


def extract_pdf_text(data: bytes, *, max_pages: int = 50) -> PDFContent | None:
    try:
        import fitz  # type: ignore
    except ImportError:
        logger.debug("pymupdf_not_installed")
        return None
    try:
        doc = fitz.open(stream=data, filetype="pdf")
        n = min(doc.page_count, max_pages)
        text = "\n\n".join(doc[i].get_text() for i in range(n)).strip()
        meta = doc.metadata or {}
        doc.close()
        if not text:
            return None
        return PDFContent(text, meta.get("title", "") or "", n, {k: str(v) for k, v in meta.items() if v})
    except Exception:  # noqa: BLE001
        logger.debug("pdf_extraction_failed")
        return None
