"""``diff_vs_reference.py interop``: byte-level and output-level agreement with the reference where nodes of both kinds
must interoperate or users would see a difference — wire frames, DHT keys, Merkle roots and proofs, PoW, signed-envelope
bytes, ledger and trust arithmetic, config keys, MCP tool schemas, and the rendered output of a local search."""
import dataclasses
import importlib
import json
import re

import diff_vs_reference as D

DOCS=[("https://docs.python.org/3/library/asyncio.html","asyncio — Asynchronous I/O","asyncio is a library to write concurrent code using the async/await syntax. asyncio is used as a foundation for multiple Python asynchronous frameworks that provide high-performance network and web-servers, database connection libraries, distributed task queues, etc.\n\nThe event loop is the core of every asyncio application. Event loops run asynchronous tasks and callbacks, perform network IO operations, and run subprocesses."),
 ("https://doc.rust-lang.org/book/ch04-01-what-is-ownership.html","What is Ownership?","Ownership is a set of rules that govern how a Rust program manages memory. All programs have to manage the way they use a computer's memory while running. Some languages have garbage collection that regularly looks for no-longer-used memory as the program runs."),
 ("https://example.org/python-history","History of Python","Python was conceived in the late 1980s by Guido van Rossum at Centrum Wiskunde & Informatica (CWI) in the Netherlands as a successor to the ABC programming language. Python 2.0 was released in 2000. Python 3.0 was released in 2008."),
 ("https://ko.example.org/python","파이썬 소개","파이썬은 배우기 쉬운 프로그래밍 언어입니다. 데이터 분석과 웹 개발에 널리 사용됩니다. 비동기 프로그래밍은 asyncio 라이브러리로 합니다.")]
QUERIES=["asyncio event loop","python","rust memory ownership","nonexistent zzz","파이썬 프로그래밍","\"quoted\" AND (python)","history of python release", "event-loop"]


def local_search_outputs(pkg):
    LS=importlib.import_module(pkg+".index.local_store").LocalStore
    Q=importlib.import_module(pkg+".search.query")
    F=importlib.import_module(pkg+".search.formatter")
    st=LS()
    for i,(u,t,x) in enumerate(DOCS):
        st.add_document(url=u,title=t,text=x,raw_html_hash=f"r{i}",text_hash=f"t{i}",language="ko" if "ko." in u else "en")
    out={}
    for q in QUERIES:
        r=Q.search_local(st,q,limit=5)
        txt=F.format_fts_results(r)
        txt=re.sub(r"\d+(\.\d+)?\s*ms","Xms",txt)
        out[q]=txt
        js=json.loads(F.format_fts_results_json(r))
        js.pop("elapsed_ms",None)
        for r_ in js.get("results", []):
            r_.pop("crawled_at", None)
        out[q+"#json"]=js
    out["suggest"]=st.suggest("Hist") if hasattr(st,"suggest") else None
    out["stats"]=st.get_stats()
    out["domains"]=st.get_top_domains(limit=5)
    return out


def both(fn):
    out = []
    for pkg in ("infomesh", "infomesh_b200"):
        try:
            out.append(("ok", D._norm(fn(lambda m, pkg=pkg: importlib.import_module(pkg + "." + m)))))
        except Exception as e:  # noqa: BLE001
            out.append(("raise", type(e).__name__ + ": " + str(e)[:120]))
    return out


CASES={}
def case(f): CASES[f.__name__]=f; return f

@case
def wire_frames(M):
    P=M("p2p.protocol")
    return [P.encode_message(P.MessageType.PING,{"t":1,"peer":"abc"}).hex(), P.encode_message(P.MessageType.SEARCH_REQUEST,{"query":"q","limit":5,"request_id":"r1"}).hex(),
            P.keyword_to_dht_key("Rust"), P.url_to_dht_key("https://e.org/a"), sorted(int(x) for x in P.MessageType if int(x) <= 100)]
@case
def wire_dataclasses(M):
    P=M("p2p.protocol")
    names=["SearchRequest","SearchResponse","PeerPointer","IndexSubmit","IndexSubmitAck","CrawlLock","ReplicateRequest","ReplicateResponse","KeyRevocationRecord","CreditProofRequest","PexRequest","PexResponse","IndexPublish","CrawlAssign","AttestationPublish"]
    return {n:[(f.name) for f in dataclasses.fields(getattr(P,n))] if hasattr(P,n) else "MISSING" for n in names}
@case
def merkle(M):
    T=M("trust.merkle")
    t=T.MerkleTree(["a","b","c","d","e"])
    p=t.get_proof(2) if hasattr(t,"get_proof") else None
    return [t.root_hash, T.MerkleTree(["x"]).root_hash, T.MerkleTree([]).root_hash if True else None, D._norm(p)]
@case
def sybil(M):
    S=M("p2p.sybil")
    return [S.compute_pow_hash(b"k"*32, 5).hex() if isinstance(S.compute_pow_hash(b"k"*32,5),bytes) else S.compute_pow_hash(b"k"*32,5), S.derive_node_id(b"k"*32,5)]
@case
def envelope(M):
    A=M("p2p.message_auth")
    fn=getattr(A,"canonical_bytes",None) or getattr(A,"_canonical_bytes",None)
    return fn(b"payload","peer",3,12.5).hex() if fn else sorted(n for n in dir(A) if "canon" in n.lower())
@case
def ledger_math(M):
    L=M("credits.ledger"); T=M("credits.types")
    led=L.CreditLedger()
    A=T.ActionType
    for a,q in [(A.CRAWL,10),(A.QUERY_PROCESS,4),(A.DOC_HOSTING,100),(A.LLM_SUMMARIZE_OWN,50)]:
        led.record_action(a,q)
    s=led.stats()
    al=led.search_allowance()
    led.spend(1000.0, reason="x")
    al2=led.search_allowance()
    return [round(s.total_earned,4), round(s.contribution_score,4), str(getattr(s.tier,"value",s.tier)), round(al.search_cost,4), al.state.value, al2.state.value, round(led.balance(),4)]
@case
def trust_scoring(M):
    S=M("trust.scoring")
    st=S.TrustStore()
    st.update_uptime("p",100.0) if hasattr(st,"update_uptime") else None
    st.record_audit("p",True); st.record_audit("p",True); st.record_audit("p",False)
    st.record_contribution("p",500.0) if hasattr(st,"record_contribution") else None
    t=st.get_trust("p") if hasattr(st,"get_trust") else st.compute_trust("p")
    return D._norm(t)
@case
def config_toml(M):
    C=M("config")
    import tempfile, pathlib
    d=pathlib.Path(tempfile.mkdtemp())
    cfg=C.Config()
    C.save_config(cfg, d/"c.toml")
    txt=(d/"c.toml").read_text()
    import re
    txt=re.sub(r'(data_dir|db_path)\s*=.*','\\1 = X',txt)
    sections=re.findall(r"^\[(\w+)\]",txt,re.M)
    keys={}
    cur=None
    for line in txt.splitlines():
        m=re.match(r"^\[(\w+)\]",line)
        if m: cur=m.group(1); keys[cur]=[]
        elif "=" in line and cur and not line.startswith("#"): keys[cur].append(line.split("=")[0].strip())
    return keys
@case
def errors_catalogue(M):
    E=M("errors")
    return {k:(v.code,v.category.value if hasattr(v.category,"value") else str(v.category)) for k,v in E.ERRORS.items()}
@case
def profiles(M):
    P=M("resources.profiles")
    return {n:D._norm(P.get_profile(n)) for n in ("minimal","balanced","contributor","dedicated")}
@case
def slo(M):
    S=M("slo")
    return D._norm(S.DEFAULT_SLOS)
@case
def plugins(M):
    P=M("plugins")
    return [h.value for h in P.HookPoint]
@case
def mcp_tools(M):
    T=M("mcp.tools")
    return [(t.name, sorted(t.inputSchema["properties"]), t.inputSchema.get("required",[])) for t in T.get_all_tools()]
@case
def seeds(M):
    S=M("crawler.seeds")
    return {c:len(S.load_seeds(c)) for c in S.CATEGORIES} if hasattr(S,"CATEGORIES") else None


def superset(ref, ours) -> bool:
    """True when ``ours`` contains everything ``ref`` has (this build adds GPU-specific entries in a few catalogues)."""
    if isinstance(ref, dict) and isinstance(ours, dict):
        return all(k in ours and superset(v, ours[k]) for k, v in ref.items())
    if isinstance(ref, list) and isinstance(ours, list):
        return len(ours) >= len(ref) and all(superset(a, b) for a, b in zip(ref, ours))
    return ref == ours


SUPERSET_OK = {"errors_catalogue", "profiles", "slo"}          # ours ⊇ reference
CURATED = {"seeds"}                                            # data lists curated independently


def run() -> int:
    bad = 0
    a, b = local_search_outputs("infomesh"), local_search_outputs("infomesh_b200")
    for k in a:
        if a[k] != b[k]:
            bad += 1
            print(f"  DIFF local search output [{k}]\n      ref : {str(a[k])[:400]}\n      ours: {str(b[k])[:400]}")
    for name, fn in CASES.items():
        r, o = both(fn)
        if name in CURATED:
            continue
        same = (r == o) or (name in SUPERSET_OK and r[0] == o[0] == "ok" and superset(r[1], o[1]))
        if not same:
            bad += 1
            print(f"  DIFF {name}\n      ref : {json.dumps(r, ensure_ascii=False, default=str)[:600]}\n      ours: {json.dumps(o, ensure_ascii=False, default=str)[:600]}")
    print(f"{len(a)} search outputs and {len(CASES)} interop cases compared, {bad} differences")
    return 1 if bad else 0
