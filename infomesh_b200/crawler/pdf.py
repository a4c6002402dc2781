"""Text out of PDF bytes, when PyMuPDF (``fitz``) is importable.

Contract (SURVEY §2.1 crawler/ "pdf"; reference infomesh/crawler/pdf.py): a URL is a PDF when its path ends in ``.pdf``
(any case, trailing slash ignored) or it carries an ``application/pdf`` hint; extraction reads at most ``max_pages`` pages,
joins them with blank lines, and yields ``None`` for an empty document, a parse failure, or a missing library -- never an
exception.  Metadata values are stringified, empty ones dropped."""
from __future__ import annotations

from contextlib import closing
from dataclasses import dataclass
from importlib import import_module

from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

_PDF_HINTS = ("application/pdf",)


@dataclass(frozen=True)
class PDFContent:
    text: str
    title: str
    page_count: int
    metadata: dict[str, str]


def is_pdf_url(url: str) -> bool:
    folded = url.casefold().rstrip("/")
    return folded.endswith(".pdf") or any(hint in folded for hint in _PDF_HINTS)


def _pymupdf():
    try:
        return import_module("fitz")
    except ImportError:
        logger.debug("pymupdf_not_installed")
        return None


def extract_pdf_text(data: bytes, *, max_pages: int = 50) -> PDFContent | None:
    engine = _pymupdf()
    if engine is None:
        return None
    try:
        with closing(engine.open(stream=data, filetype="pdf")) as document:
            pages = min(max_pages, document.page_count)
            body = "\n\n".join(document[number].get_text() for number in range(pages)).strip()
            info = dict(document.metadata or {})
    except Exception:  # noqa: BLE001 -- damaged or encrypted file
        logger.debug("pdf_extraction_failed")
        return None
    if not body:
        return None
    return PDFContent(text=body, title=info.get("title") or "", page_count=pages,
                      metadata={key: str(value) for key, value in info.items() if value})
