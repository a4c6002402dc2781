"""``diff_vs_reference.py http``: every route of the local admin API through FastAPI's test client on both sides; JSON bodies must
contain everything the reference returns (volatile fields such as uptimes are ignored)."""
import importlib
import json
import re
import tempfile
import time
from dataclasses import replace

from diff_interop import DOCS, superset


def side(pkg):
    M=lambda m: importlib.import_module(pkg+"."+m)
    C=M("config"); A=M("api.local_api"); LS=M("index.local_store").LocalStore
    d=tempfile.mkdtemp(); P=__import__("pathlib").Path
    base=C.Config()
    cfg=replace(base,node=replace(base.node,data_dir=P(d)),index=replace(base.index,db_path=P(d)/"index.db",vector_search=False),storage=replace(base.storage,compression_enabled=False))
    st=LS(db_path=cfg.index.db_path)
    for i,(u,t,x) in enumerate(DOCS): st.add_document(url=u,title=t,text=x,raw_html_hash=f"r{i}",text_hash=f"t{i}",language="en")
    st.close()
    from fastapi.testclient import TestClient

    c=TestClient(A.create_admin_app(cfg))
    out={}
    for path in ("/health","/health?detail=1","/readiness","/status","/index/stats","/index/compression","/credits/balance","/network/peers","/analytics","/analytics/tools","/search?q=python&limit=2","/search","/config","/openapi-spec"):
        r=c.get(path)
        try: body=r.json()
        except Exception: body=r.text[:200]
        out[path]=(r.status_code, body)
        time.sleep(0.12)
    r=c.post("/config/reload"); out["POST /config/reload"]=(r.status_code,r.json())
    m=c.get("/metrics"); out["/metrics"]=(m.status_code, sorted(set(re.findall(r"^(infomesh_[a-z_]+)", m.text if "json" not in m.headers.get("content-type","") else m.json().get("metrics",""), re.M))))
    out["headers"]=sorted(k for k in c.get("/health").headers if k.lower().startswith(("x-","content-security")))
    return out


def strip(o):
    if isinstance(o,dict): return {k:strip(v) for k,v in o.items() if k not in ("uptime_seconds","uptime_human","elapsed_ms","uptime_s","disk_free_gb","memory_pct","version","data_dir","db_path","db_size_mb","avg_latency_ms","avg_doc_kb")}
    if isinstance(o,list): return [strip(x) for x in o]
    return o


def run() -> int:
    from fastapi.testclient import TestClient  # noqa: F401 — fail early with a clear error when fastapi is missing

    a, b = side("infomesh"), side("infomesh_b200")
    bad = 0
    for k in a:
        ra, rb = a[k], b[k]
        if k == "/metrics":                                  # metric names: ours must include the reference's
            ok = ra[0] == rb[0] and set(ra[1]) <= set(rb[1])
        elif k == "/openapi-spec":                           # richer here; the reference's paths and schemas must be present
            ok = ra[0] == rb[0] and set(ra[1]["paths"]) <= set(rb[1]["paths"]) and set(ra[1]["components"]["schemas"]) <= set(rb[1]["components"]["schemas"])
        else:
            ok = ra[0] == rb[0] and (strip(ra[1]) == strip(rb[1]) or superset(strip(ra[1]), strip(rb[1])))
        if not ok:
            bad += 1
            print(f"  DIFF http {k}\n      ref : {json.dumps(strip(ra), default=str)[:500]}\n      ours: {json.dumps(strip(rb), default=str)[:500]}")
    print(f"{len(a)} admin API responses compared, {bad} differences")
    return 1 if bad else 0
