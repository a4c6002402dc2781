"""Crawl one page (robots.txt, SSRF checks, dedup and link extraction included) and search it."""
import sys
import tempfile

from infomesh_b200.sdk import InfoMeshClient

url = sys.argv[1] if len(sys.argv) > 1 else "https://docs.python.org/3/library/asyncio.html"
with tempfile.TemporaryDirectory() as d, InfoMeshClient(d, {"crawl.politeness_delay": 0.5}) as client:
    res = client.crawl(url)
    print("crawl:", res)
    if res.success:
        for hit in client.search(" ".join(res.title.split()[:2]) or "asyncio", limit=5):
            print(f"{hit.score:.3f}  {hit.title}  <{hit.url}>")
