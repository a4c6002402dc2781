"""``diff_vs_reference.py mcp``: the text every MCP tool returns, reference handlers vs this build, over the same four documents.
JSON answers must contain everything the reference returns (this build adds fields such as ``gpu`` or ``sources``)."""
import asyncio
import importlib
import inspect
import json
import re

from diff_interop import DOCS, superset

# deliberate behaviour differences, each better for the caller:
#  * structured answers are extracted from snippets with the <b> highlight tags removed (the reference matches "<b>python</b>" literally)
#  * fetch_page serves a cached page and validates the URL even when no crawl worker is attached
DELIBERATE = {"web_search json", "fetch cached", "fetch blocked"}


def side(pkg):
    M=lambda m: importlib.import_module(pkg+"."+m)
    H=M("mcp.handlers"); LS=M("index.local_store").LocalStore; C=M("config"); S=M("mcp.session"); QC=M("search.cache").QueryCache
    st=LS()
    for i,(u,t,x) in enumerate(DOCS):
        st.add_document(url=u,title=t,text=x,raw_html_hash=f"r{i}",text_hash=f"t{i}",language="ko" if "ko." in u else "en")
    cfg=C.Config()
    deps=dict(config=cfg,store=st,vector_store=None,distributed_index=None,p2p_node=None,link_graph=None,ledger=None,llm_backend=None,query_cache=QC(),sessions=S.SessionStore(),analytics=S.AnalyticsTracker())
    out={}
    def run(name, fn, *a, **k):
        try:
            sig=inspect.signature(fn).parameters
            kk={n:v for n,v in k.items() if n in sig or any(p.kind==p.VAR_KEYWORD for p in sig.values())}
            r=fn(*a,**kk)
            if inspect.iscoroutine(r): r=asyncio.run(r)
            txt=r[0].text
        except Exception as e:
            txt="RAISE "+type(e).__name__+": "+str(e)[:200]
        txt=re.sub(r"\d+(\.\d+)?\s*ms","Xms",txt); txt=re.sub(r'"elapsed_ms": [\d.]+','"elapsed_ms": 0',txt); txt=re.sub(r'"crawled_at": [\d.]+','"crawled_at": 0',txt)
        txt=re.sub(r'"timestamp": [\d.]+','"timestamp": 0',txt); txt=re.sub(r'"version": "[^"]+"','"version": "X"',txt); txt=re.sub(r"crawled_at: [\d.\-T:+Z ]+","crawled_at: X",txt)
        txt = re.sub(r'"analytics": \{[^}]*\}', '"analytics": {}', txt)      # counters depend on what ran before
        out[name]=txt
    run("web_search text", H.handle_web_search, {"query":"asyncio event loop","local_only":True}, **deps)
    run("web_search json", H.handle_web_search, {"query":"python","local_only":True,"top_k":3,"answer_mode":"structured"}, **deps)
    run("web_search explain", H.handle_web_search, {"query":"rust memory","local_only":True,"explain":True}, **deps)
    run("web_search empty", H.handle_web_search, {"query":"zzzqqq","local_only":True}, **deps)
    run("web_search noquery", H.handle_web_search, {}, **deps)
    run("web_search filters", H.handle_web_search, {"query":"python","local_only":True,"domain_allowlist":["example.org"],"language":"en"}, **deps)
    run("search_local", H.handle_search, "search_local", {"query":"python","limit":2,"format":"json"}, **deps)
    run("fetch cached", H.handle_fetch, {"url":DOCS[2][0]}, config=cfg, store=st, worker=None, vector_store=None, link_graph=None, analytics=deps["analytics"], feedback_store=None, last_search_query="")
    run("fetch blocked", H.handle_fetch, {"url":"http://127.0.0.1/x"}, config=cfg, store=st, worker=None, vector_store=None, link_graph=None, analytics=deps["analytics"], feedback_store=None, last_search_query="")
    run("fetch missing", H.handle_fetch, {"url":"https://example.org/none"}, config=cfg, store=st, worker=None, vector_store=None, link_graph=None, analytics=deps["analytics"], feedback_store=None, last_search_query="")
    run("suggest", H.handle_suggest, {"prefix":"Hist"}, store=st)
    run("explain", H.handle_explain, {"query":"python history"}, store=st, link_graph=None)
    run("fact_check", H.handle_fact_check, {"claim":"Python 3.0 was released in 2008"}, store=st, link_graph=None)
    run("search_rag", H.handle_search_rag, {"query":"asyncio event loop"}, store=st, link_graph=None, analytics=deps["analytics"], ledger=None)
    run("extract_answer", H.handle_extract_answer, {"query":"who created python"}, store=st, link_graph=None, ledger=None)
    run("batch", H.handle_batch, {"queries":["python","rust"]}, store=st, link_graph=None, ledger=None, analytics=deps["analytics"])
    run("status", H.handle_status, {}, store=st, vector_store=None, link_graph=None, ledger=None, scheduler=None, p2p_node=None, distributed_index=None, analytics=deps["analytics"])
    run("remove", H.handle_remove_url, {"url":DOCS[3][0]}, store=st)
    run("crawl blocked", H.handle_crawl, {"url":"http://10.0.0.1/"}, config=cfg, store=st, worker=None, vector_store=None, distributed_index=None, p2p_node=None, link_graph=None, analytics=deps["analytics"], webhooks=S.WebhookRegistry())
    return out


def run() -> int:
    a, b = side("infomesh"), side("infomesh_b200")
    bad = 0
    for k in a:
        if a[k] == b[k] or k in DELIBERATE:
            continue
        try:
            ok = superset(json.loads(a[k]), json.loads(b[k]))
        except ValueError:
            ok = False
        if not ok:
            bad += 1
            print(f"  DIFF mcp tool output [{k}]\n      ref : {a[k][:500]}\n      ours: {b[k][:500]}")
    print(f"{len(a)} MCP tool outputs compared ({len(DELIBERATE)} deliberate differences), {bad} unexpected differences")
    return 1 if bad else 0
