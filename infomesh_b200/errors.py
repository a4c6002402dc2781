"""Structured error catalogue: ``INFOMESH_E001`` .. ``E020`` (plus the GPU-plane codes E101+ this build adds), each
with a category, message, resolution hint and HTTP status (reference infomesh/errors.py:12-197)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import StrEnum


class ErrorCategory(StrEnum):
    AUTH = "AUTH"
    SEARCH = "SEARCH"
    CRAWL = "CRAWL"
    INDEX = "INDEX"
    NETWORK = "NETWORK"
    CONFIG = "CONFIG"
    RESOURCE = "RESOURCE"
    SECURITY = "SECURITY"
    GPU = "GPU"


@dataclass(frozen=True)
class InfoMeshError:
    code: str
    category: ErrorCategory
    message: str
    resolution: str
    http_status: int = 400

    def to_dict(self) -> dict[str, object]:
        return {"error": {"code": self.code, "category": self.category.value, "message": self.message,
                          "resolution": self.resolution}}

    def format(self) -> str:
        return f"Error [{self.code}]: {self.message}\nResolution: {self.resolution}"


_C = ErrorCategory
_TABLE: tuple[tuple[str, ErrorCategory, str, str, int], ...] = (
    ("E001", _C.AUTH, "Invalid or missing API key", "Set INFOMESH_API_KEY env var or pass api_key parameter", 401),
    ("E002", _C.AUTH, "Insufficient role permissions", "Contact admin to assign appropriate role (admin/reader/crawler)", 403),
    ("E003", _C.SEARCH, "Query must be a non-empty string", "Provide a non-empty query parameter", 400),
    ("E004", _C.SEARCH, "Query exceeds maximum length (1000 chars)", "Shorten query to under 1000 characters", 400),
    ("E005", _C.CRAWL, "URL blocked for security reasons (SSRF)", "Only http:// and https:// public URLs are allowed", 400),
    ("E006", _C.CRAWL, "Crawl rate limit exceeded (60 URLs/hour)", "Wait before submitting more URLs or increase rate limit in config", 429),
    ("E007", _C.CRAWL, "robots.txt disallows crawling this URL",
     "This URL is blocked by the site's robots.txt. Choose a different URL.", 400),
    ("E008", _C.INDEX, "Document exceeds maximum size", "Reduce document size or increase max_doc_size_kb in config", 400),
    ("E009", _C.NETWORK, "No P2P peers connected", "Check network connectivity and bootstrap node configuration", 400),
    ("E010", _C.RESOURCE, "Insufficient disk space", "Free up disk space (minimum 500MB required)", 400),
    ("E011", _C.CONFIG, "Invalid configuration value",
     "Check config.toml for valid values. Run 'infomesh config show' to review.", 400),
    ("E012", _C.SEARCH, "Batch search exceeds maximum queries (10)", "Reduce batch size to 10 or fewer queries", 400),
    ("E013", _C.SECURITY, "IP address blocked by access control", "Contact admin to add IP to allowlist", 403),
    ("E014", _C.AUTH, "JWT token expired or invalid", "Obtain a new JWT token and retry", 401),
    ("E015", _C.CRAWL, "Crawler worker not available", "Start the node with crawl capability or use role=full", 400),
    ("E016", _C.SEARCH, "Vector search not available", "Enable [index] vector_search and make sure the encoder can load", 400),
    ("E017", _C.RESOURCE, "Node in defensive mode (overloaded)", "Reduce load or wait for resource governor to recover", 400),
    ("E018", _C.NETWORK, "Port already in use", "Stop the existing process or use a different port", 400),
    ("E019", _C.CRAWL, "Paywall detected", "This content is behind a paywall and cannot be retrieved", 400),
    ("E020", _C.AUTH, "Rate limit exceeded for API key", "Wait for rate limit window to reset or contact admin", 429),
    ("E101", _C.GPU, "Native kernel library not loadable", "Run `python -m infomesh_b200.build` (needs nvcc, sm_100a)", 500),
    ("E102", _C.GPU, "No sm_100a device visible", "Run on a B200 or set [gpu] enabled = false to use the CPU search plane", 500),
    ("E103", _C.GPU, "GPU index does not fit in device memory", "Lower [gpu] memory_fraction or shard over more GPUs", 507),
    ("E104", _C.GPU, "Peer-memory exchange timed out", "A rank stalled or died: check `nvidia-smi` and restart the job", 500),
)
ERRORS: dict[str, InfoMeshError] = {k: InfoMeshError(f"INFOMESH_{k}", c, m, r, s) for k, c, m, r, s in _TABLE}


def get_error(code: str) -> InfoMeshError | None:
    return ERRORS.get(code)


def format_error(code: str) -> str:
    err = ERRORS.get(code)
    return err.format() if err else f"Unknown error: {code}"
