"""Extended search: bounded-parallel batch search (<= 50 queries, 5 at a time), a TTL cache for generated summaries,
keyword-level query translation for nine languages (reference infomesh/search/extended.py:21-303), and the device
batch path: :func:`batch_search_gpu` packs up to 128 queries into one pass of the hybrid GPU pipeline."""
from __future__ import annotations

import asyncio
import hashlib
import time
from dataclasses import dataclass, field
from typing import Any, Awaitable, Callable

MAX_BATCH_QUERIES = 50


@dataclass
class BatchQuery:
    query: str
    top_k: int = 5
    language: str | None = None


@dataclass
class BatchResult:
    query: str
    results: list[dict[str, object]] = field(default_factory=list)
    elapsed_ms: float = 0.0
    error: str | None = None


@dataclass
class BatchSearchResponse:
    results: list[BatchResult] = field(default_factory=list)
    total_elapsed_ms: float = 0.0
    total_queries: int = 0


async def batch_search(queries: list[BatchQuery], search_fn: Callable[[str, int, str | None], Awaitable[Any]], *,
                       max_parallel: int = 5) -> BatchSearchResponse:
    t0 = time.monotonic()
    gate = asyncio.Semaphore(max_parallel)

    async def one(q: BatchQuery) -> BatchResult:
        async with gate:
            s = time.monotonic()
            try:
                res = await search_fn(q.query, q.top_k, q.language)
                return BatchResult(q.query, res if isinstance(res, list) else [], (time.monotonic() - s) * 1000)
            except Exception as exc:  # noqa: BLE001 — one bad query must not fail the batch
                return BatchResult(q.query, error=str(exc), elapsed_ms=(time.monotonic() - s) * 1000)

    done = await asyncio.gather(*(one(q) for q in queries[:MAX_BATCH_QUERIES]))
    return BatchSearchResponse(list(done), (time.monotonic() - t0) * 1000, len(queries))


def batch_search_gpu(queries: list[BatchQuery], gpu_index: Any) -> BatchSearchResponse:
    """All queries in ONE device pass (``engine.gpu_index.GpuSearchIndex.search_many``): encoder, dense + BM25
    retrieval, fusion and rerank are batched over the query dimension instead of fanned out as tasks."""
    t0 = time.monotonic()
    qs = queries[:128]
    try:
        hits = gpu_index.search_many([q.query for q in qs], k=max((q.top_k for q in qs), default=5))
        out = [BatchResult(q.query, list(h)[:q.top_k]) for q, h in zip(qs, hits)]
    except Exception as exc:  # noqa: BLE001
        out = [BatchResult(q.query, error=str(exc)) for q in qs]
    ms = (time.monotonic() - t0) * 1000
    for r in out:
        r.elapsed_ms = ms
    return BatchSearchResponse(out, ms, len(queries))


@dataclass
class CachedSummary:
    query_hash: str
    summary: str
    sources: list[str]
    created_at: float
    expires_at: float


class SummaryCache:
    def __init__(self, max_entries: int = 500, ttl_seconds: float = 3600):
        self._cache: dict[str, CachedSummary] = {}
        self._max, self._ttl = max_entries, ttl_seconds

    @staticmethod
    def _hash(query: str) -> str:
        return hashlib.sha256(query.strip().lower().encode()).hexdigest()[:16]

    def get(self, query: str) -> CachedSummary | None:
        h = self._hash(query)
        e = self._cache.get(h)
        if e is None:
            return None
        if time.time() < e.expires_at:
            return e
        del self._cache[h]
        return None

    def put(self, query: str, summary: str, sources: list[str]) -> None:
        h = self._hash(query)
        if h not in self._cache and len(self._cache) >= self._max:
            del self._cache[min(self._cache, key=lambda k: self._cache[k].created_at)]
        now = time.time()
        self._cache[h] = CachedSummary(h, summary, list(sources), now, now + self._ttl)

    @property
    def size(self) -> int:
        return len(self._cache)


def _table(words: str, english: str) -> dict[str, str]:
    return dict(zip(words.split("|"), english.split("|")))


_EN9 = "install|error|configuration|search|file|server|database|network|security"
_TRANSLATIONS: dict[str, dict[str, str]] = {
    "ko": _table("설치|사용법|오류|설정|검색|파일|서버|데이터베이스|네트워크|보안|성능|테스트|배포|업데이트|삭제",
                 "install|usage|error|configuration|search|file|server|database|network|security|performance|test|deploy|update|delete"),
    "ja": _table("インストール|エラー|設定|検索|ファイル|サーバー|データベース|ネットワーク|セキュリティ", _EN9),
    "zh": _table("安装|错误|配置|搜索|文件|服务器|数据库|网络|安全|性能", _EN9 + "|performance"),
    "ar": _table("تثبيت|خطأ|إعدادات|بحث|ملف|خادم|قاعدة بيانات|شبكة|أمان", _EN9),
    "hi": _table("स्थापित|त्रुटि|सेटिंग|खोज|फ़ाइल|सर्वर|डेटाबेस|नेटवर्क|सुरक्षा", _EN9),
    "th": _table("ติดตั้ง|ข้อผิดพลาด|การตั้งค่า|ค้นหา|ไฟล์|เซิร์ฟเวอร์|ฐานข้อมูล|เครือข่าย", _EN9.rsplit("|", 1)[0]),
    "tr": _table("kurulum|hata|ayarlar|arama|dosya|sunucu|veritabanı|ağ|güvenlik", _EN9),
    "vi": _table("cài đặt|lỗi|cấu hình|tìm kiếm|tập tin|máy chủ|cơ sở dữ liệu|mạng|bảo mật", _EN9),
    "id": _table("instalasi|kesalahan|pengaturan|pencarian|berkas|server|basis data|jaringan|keamanan", _EN9),
}


def translate_query_keywords(query: str, source_lang: str) -> list[str]:
    """English equivalents of the known terms that occur in the query (added alongside the original terms)."""
    return [eng for term, eng in _TRANSLATIONS.get(source_lang, {}).items() if term in query]
