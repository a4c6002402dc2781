"""Python SDK: use InfoMesh as a library."""
from infomesh_b200.sdk.client import CrawlResult, InfoMeshClient, NetworkInfo, SearchResult

__all__ = ["InfoMeshClient", "SearchResult", "CrawlResult", "NetworkInfo"]
