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


# ---- stateful scenarios: the same script drives both implementations through the reference's method names ----
@case
def feedback_store(M):
    F = M("search.feedback").FeedbackStore()
    for _ in range(3):
        F.record_fetch("tensor memory", "https://a.org/1", 1)
    F.record_skip("tensor memory", ["https://b.org/2", "https://b.org/3"])
    F.record_citation("tensor memory", "https://a.org/1")
    F.record_reformulation("tensor memory")
    st = F.get_url_stats("https://a.org/1")
    out = [round(F.get_boost("https://a.org/1"), 6), round(F.get_boost("https://b.org/2"), 6), F.get_boost("https://none/"), F.signal_count(),
           [(u.url, round(u.boost_score, 6), u.fetch_count, u.skip_count, u.cite_count) for u in F.top_boosted_urls(5)],
           (st.fetch_count, st.skip_count, st.cite_count) if st else None, len(F.hash_query("Tensor Memory ")), F.hash_query("a") == F.hash_query("A ")]
    F.close()
    return out


@case
def subnet_limiter_and_profiles(M):
    S = M("p2p.sybil").SubnetLimiter()
    res = [S.can_add("10.1.2.3", 0), S.add("10.1.2.3", "p1", 0), S.add("10.1.2.4", "p2", 0), S.add("10.1.2.5", "p3", 0), S.can_add("10.1.2.6", 0),
           S.add("10.1.2.6", "p4", 0), S.can_add("10.1.2.6", 1), S.get_subnet_counts(0), S.total_nodes]
    S.remove("10.1.2.3", "p1", 0)
    res.append(S.can_add("10.1.2.6", 0))
    T = M("p2p.peer_profile").PeerProfileTracker()
    for ms in (100, 300, 200):
        T.record("a", ms)
    T.record("b", 2000, success=False)
    T.record("c", 50)
    pa = T.get("a")
    res += [round(pa.avg_latency_ms, 4), round(T.adaptive_timeout("a", base_ms=5000), 3), T.adaptive_timeout("zz"), T.known_peers,
            T.rank_by_latency(["a", "b", "c"], diversity=False), round(T.get("b").success_rate, 3)]
    return res


@case
def simhash_index_and_dedup(M):
    SH = M("crawler.simhash")
    idx = SH.SimHashIndex(max_entries=3)
    fp = SH.simhash("the quick brown fox jumps over the lazy dog near the river bank today")
    idx.add(1, fp), idx.add(2, fp ^ 0b101), idx.add(3, fp ^ (2 ** 40 - 1))
    res = [sorted(idx.find_near_duplicates(fp)), sorted(idx.find_near_duplicates(fp, threshold=0)), idx.size]
    idx.add(4, 12345)                                   # over capacity: the oldest entry leaves
    res += [idx.size, sorted(idx.find_near_duplicates(fp))]
    DD = M("crawler.dedup").DeduplicatorDB()
    text = "Thread block clusters let several CTAs share distributed shared memory and multicast TMA loads across the cluster. " * 3
    res += [DD.is_url_seen("https://e.org/a?utm_source=x"), DD.is_content_seen("h1"), DD.is_near_duplicate(text)]
    DD.mark_seen("https://e.org/a", "h1", text)
    res += [DD.is_url_seen("https://E.org/a/?utm_source=x#frag"), DD.is_content_seen("h1"), DD.is_near_duplicate(text + " extra"), DD.is_near_duplicate("completely different words about cooking pasta with tomato sauce and basil leaves")]
    DD.close()
    return res


@case
def bloom_sessions_webhooks(M):
    B = M("scalability").BloomFilter(capacity=1000, fp_rate=0.01)
    for i in range(200):
        B.add(f"url-{i}")
    res = [all(f"url-{i}" in B for i in range(200)), sum(f"other-{i}" in B for i in range(500)) < 25, B.size_bytes]
    X = M("mcp.session")
    st = X.SessionStore(max_size=2)
    a = st.get_or_create("s1")
    st.get_or_create("s2"), st.get_or_create("s3")
    res.append(st.get_or_create("s1") is a)             # s1 was evicted by the size cap
    an = X.AnalyticsTracker()
    an.record_search(10.0), an.record_search(30.0), an.record_crawl(), an.record_fetch()
    d = an.to_dict()
    res.append({k: d[k] for k in ("total_searches", "total_crawls", "total_fetches", "avg_latency_ms") if k in d})
    wh = X.WebhookRegistry(max_registrations=2)
    res += [wh.register("https://hooks.example.org/a"), wh.register("http://127.0.0.1/x"), wh.register("https://hooks.example.org/a"),
            wh.register("https://hooks.example.org/b"), wh.register("https://hooks.example.org/c"), wh.unregister("https://hooks.example.org/a"), sorted(wh.urls)]
    return res


@case
def quality_and_related(M):
    Qm = M("search.quality")
    ab = Qm.ABTest("rerank")
    ab.compare("q1", [3, 2, 0], [0, 2, 3]), ab.compare("q2", [1, 0], [1, 0]), ab.compare("q3", [0, 1], [1, 0])
    c = Qm.QueryIntentClassifier()
    res = [D._norm(ab.summary()), [c.classify(q) for q in ("how to install cuda", "facebook login", "buy rtx 5090", "latest ai news today", "python")],
           D._norm(c.classify_with_confidence("how do I write a kernel"))]
    R = M("search.nlp").RelatedSearchTracker()
    for q in ("python asyncio", "python threading", "python asyncio tutorial", "rust ownership"):
        R.record(q)
    res.append(R.related("python asyncio", limit=3))
    return res


@case
def reputation_and_farming(M):
    T = M("trust.reputation").LLMReputationTracker()
    for q in (0.9, 0.8, 1.0, 0.7, 0.9, 0.95):
        T.record_quality("good", q)
    for q in (0.1, 0.2, 0.1, 0.3, 0.2):
        T.record_quality("bad", q)
    T.record_quality("new", 0.9)
    g, b = T.get_reputation("good"), T.get_reputation("bad")
    res = [round(g.ema_quality, 4), str(getattr(g.grade, "value", g.grade)), round(b.ema_quality, 4), str(getattr(b.grade, "value", b.grade)),
           round(T.get_quality_score("nobody"), 4), [p.peer_id for p in T.top_peers(5)], [p.peer_id for p in T.list_peers()]]
    F = M("credits.farming").FarmingDetector()
    t0 = 1_000_000.0
    F.register_node("n1", now=t0)
    res += [F.is_on_probation("n1", now=t0 + 3600), F.is_on_probation("n1", now=t0 + 25 * 3600), round(F.probation_remaining("n1", now=t0 + 3600), 1)]
    for i in range(40):
        F.log_action("n1", "crawl", now=t0 + 30 * 3600 + i * 60.0)          # metronome-regular: a bot signature
    now = t0 + 30 * 3600 + 40 * 60
    res += [F.actions_in_last_hour("n1", "crawl", now=now), F.detect_regular_intervals("n1", "crawl", now=now), F.detect_burst("n1", "crawl", now=now),
            F.is_rate_limited("n1", "crawl", now=now), F.is_blocked("n1")]
    v = F.check("n1", "crawl", now=now)
    res.append(D._norm(v))
    return res


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
