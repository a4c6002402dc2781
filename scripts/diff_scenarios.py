"""Stateful differential scenarios: the same sequence of public-API calls is driven through the reference package and
through this one, and the observable results are compared (tests/test_differential_cases.py, one pytest case each).

A scenario is ``name -> fn(pkg, tmp_path)`` returning plain data; ``pkg`` is ``"infomesh"`` or ``"infomesh_b200"``.
Wall-clock fields are stripped or replaced by explicit ``now=`` arguments so that both sides are deterministic."""
from __future__ import annotations

import asyncio
import importlib
from unittest import mock


def _m(pkg: str, name: str):
    return importlib.import_module(f"{pkg}.{name}")


# ----------------------------------------------------------------------------- crawler
def scheduler_bookkeeping(pkg, tmp):
    S = _m(pkg, "crawler.scheduler")

    async def go():
        s = S.Scheduler(politeness_delay=0.0, urls_per_hour=0, pending_per_domain=2, max_depth=2)
        added = [await s.add_url(u, d) for u, d in (("https://a.example/1", 0), ("https://a.example/2", 1), ("https://a.example/3", 0),
                                                      ("https://b.example/x", 3), ("https://b.example/y", 2))]
        first = await s.get_url()
        s.mark_error(first[0])
        second = await s.get_url()
        s.mark_done(second[0])
        s.mark_done("https://unknown.example/")
        s.set_crawl_delay("slow.example", 600)
        s.set_crawl_delay("fine.example", 2.5)
        room_again = [await s.add_url("https://a.example/4"), await s.add_url("https://a.example/5"), await s.add_url("https://a.example/6")]
        st = s._domains["a.example"]
        return {"added": added, "order": [first, second], "pending": s.pending_count, "room_again": room_again,
                "a": (st.pending_count, st.error_count), "slow": s._domains["slow.example"].crawl_delay, "fine": s._domains["fine.example"].crawl_delay}

    return asyncio.run(go())


def freshness_queue(pkg, tmp):
    F = _m(pkg, "crawler.freshness")
    q = F.PriorityRecrawlQueue(max_size=5)
    T = F.RecrawlTrigger
    plan = [("s1", T.SCHEDULED, 10), ("p1", T.PEER_ANNOUNCE, 11), ("u1", T.USER_REQUEST, 12), ("r1", T.RSS_UPDATE, 13), ("u1", T.RSS_UPDATE, 14),
            ("c1", T.CONTENT_CHANGE, 15), ("overflow", T.USER_REQUEST, 16)]
    accepted = [q.enqueue(u, t, now=float(n)) for u, t, n in plan]
    peek = q.peek().url
    q.discard("r1")
    order = []
    while (item := q.dequeue()) is not None:
        order.append((item.url, item.trigger.value, item.priority))
    hdr = F.ConditionalHeaders.from_response_headers({"ETag": 'W/"x"', "Last-Modified": "Mon, 01 Jan 2024 00:00:00 GMT", "X": "y"})
    return {"accepted": accepted, "peek": peek, "order": order, "size": q.size, "enq": q.total_enqueued, "deq": q.total_dequeued,
            "tiers": [F.classify_freshness(1000.0, now=1000.0 + a).value for a in (0, 3600, 3601, 86400, 86401, 604800, 604801)],
            "headers": hdr.to_request_headers(), "empty": F.ConditionalHeaders().to_request_headers()}


OPML = ('<?xml version="1.0"?><opml version="2.0"><body><outline text="Tech"><outline type="rss" text="Blog A" xmlUrl="https://a.example/feed"/>'
        '<outline type="rss" title="Blog B" xmlUrl="https://b.example/rss" htmlUrl="https://b.example"/></outline>'
        '<outline text="dup" xmlUrl="https://a.example/feed"/><outline text="no url"/></body></opml>')
RSS = ('<?xml version="1.0"?><rss><channel><title>F</title><item><title>One</title><link>https://a.example/1</link></item>'
       '<item><title>Two</title><link>https://a.example/2</link></item><item><title>Dup</title><link>https://a.example/1</link></item></channel></rss>')


def feed_monitor(pkg, tmp):
    FM = _m(pkg, "crawler.feed_monitor")
    m = FM.FeedMonitor()
    parsed = [(f.url, f.label, f.priority.value) for f in FM.parse_opml(OPML)]
    n_added = (m.add_feeds_from_opml(OPML), m.add_feeds_from_opml(OPML))
    hi = m.add_feed("https://c.example/feed", priority=FM.FeedPriority.CRITICAL, poll_interval=10, label="crit")
    lo = m.add_feed("https://d.example/feed", priority=FM.FeedPriority.LOW)
    overflow = None
    due0 = [f.url for f in m.get_due_feeds(now=100.0)]
    for f in m.feeds:
        f.last_poll_at = 100.0
    due = {t: [f.url for f in m.get_due_feeds(now=float(t))] for t in (105, 111, 401, 1001, 3701)}
    m.mark_url_seen("https://a.example/2")
    up1 = m.process_feed_response("https://a.example/feed", RSS, now=200.0)
    up2 = m.process_feed_response("https://a.example/feed", RSS, now=300.0)
    bad = m.process_feed_response("https://nobody.example/feed", RSS, now=300.0)
    again = m.add_feed("https://d.example/feed", priority=FM.FeedPriority.HIGH, label="relabel")
    st = m.stats
    a = next(f for f in m.feeds if f.url == "https://a.example/feed")
    return {"parsed": parsed, "added": n_added, "overflow": overflow, "due0": due0, "due": due, "new1": up1.new_urls, "new2": up2.new_urls,
            "err": bad.error, "same_obj": again is lo, "lo": (lo.priority.value, lo.label, lo.effective_interval), "hi": hi.effective_interval,
            "stats": (st.total_feeds, st.total_polls, st.total_new_urls, st.total_errors, dict(sorted(st.feeds_by_priority.items()))),
            "a": (a.last_poll_at, a.items_discovered, a.last_item_url, a.error_count), "removed": (m.remove_feed("https://c.example/feed"), m.remove_feed("nope"))}


def crawl_intelligence(pkg, tmp):
    I = _m(pkg, "crawler.intelligence")
    rc = I.RobotsCache(ttl_seconds=1000)
    rc.put("a.example", True, 1.5, ["https://a.example/sitemap.xml"])
    rc.put("b.example", False)
    exported = sorted(({k: v for k, v in e.items() if k != "cached_at"} for e in rc.export_for_dht()), key=lambda e: e["domain"])
    other = I.RobotsCache()
    other.put("b.example", True, 9.0)
    taken = other.import_from_dht(rc.export_for_dht() + [{"domain": "", "allowed": True}, {"domain": "c.example", "crawl_delay": "2.5", "sitemaps": "nope"},
                                                         {"domain": "d.example", "crawl_delay": "7", "allowed": 0}, {"domain": "e.example", "crawl_delay": None}])
    # (a non-numeric crawl_delay string makes the reference raise ValueError; this repo treats it as 0 -- deliberately not compared)
    view = {d: ((e.allowed, e.crawl_delay, e.sitemaps) if (e := other.get(d)) else None) for d in ("a.example", "b.example", "c.example", "d.example", "e.example", "zzz")}
    expired = I.RobotsCache(ttl_seconds=-1)
    expired.put("x.example", True)
    tuner = I.CrawlSpeedTuner(base_delay=1.0, min_delay=0.5, max_delay=3.0)
    trace = []
    import types

    import psutil

    for cpu, mem in ((95, 10), (95, 10), (95, 95), (95, 95), (75, 10), (10, 85), (50, 60), (10, 10), (10, 10), (10, 10), (10, 10), (29.9, 49.9), (30, 50)):
        with mock.patch.object(psutil, "cpu_percent", lambda interval=None, _c=cpu: _c), \
                mock.patch.object(psutil, "virtual_memory", lambda _m=mem: types.SimpleNamespace(percent=_m)):
            s = tuner.adjust()
        trace.append((s.current_delay, s.adjustment_reason.split(" (")[0], s.cpu_usage, s.memory_usage))
    alts = I.extract_image_alt_texts('<p><img src="a.png" alt="A chart of bandwidth"><img alt=""><img src=b alt="x"><img alt="Logo"><IMG ALT="Diagram of the cluster" src=c>'
                                     '<img alt="abc"><img alt="abcd"><img alt=" padded text "><img alt="ok!"></p>')
    # (repeated alt texts are reported once here and every time by the reference; entities are decoded here: not compared)
    return {"exported": exported, "taken": taken, "view": view, "size": other.size, "expired_get": expired.get("x.example"), "expired_size": expired.size,
            "cleanup": I.RobotsCache(ttl_seconds=-1).cleanup(), "trace": trace, "delay": round(tuner.current_delay, 6), "alts": alts}


def url_assignment(pkg, tmp):
    U = _m(pkg, "crawler.url_assigner")
    a = U.UrlAssigner("peer-local")
    for p in ("peer-b", "peer-c", "peer-d", "peer-b"):
        a.add_peer(p)
    a.remove_peer("peer-local")
    a.remove_peer("peer-d")
    urls = [f"https://site{i}.example/page/{i * 7}" for i in range(40)]
    asg = a.assign("https://x.example/", depth=2)
    return {"known": a.known_peers, "owners": [a.closest_peer(u) for u in urls], "local": a.filter_local_urls(urls),
            "assign": (asg.url, asg.depth, asg.priority, asg.assigner_peer_id)}


def pdf_and_structured(pkg, tmp):
    P, S = _m(pkg, "crawler.pdf"), _m(pkg, "crawler.structured")
    html = ('<html><head><meta name="description" content=" A page about GPUs "><meta name="keywords" content="gpu, cuda , ,tensor">'
            '<meta property="og:title" content="OG"><meta name="og:type" content="article"><meta property="og:title" content="OG2">'
            '<script type="application/ld+json">[{"@type":"A"},{"@type":"B"}]</script><script type="application/ld+json">{broken</script>'
            '<script type="application/ld+json">{"@type":"C"}</script></head></html>')
    sd = S.extract_structured_data(html)
    return {"is_pdf": [P.is_pdf_url(u) for u in ("https://x.org/a.PDF", "https://x.org/a.pdf/", "https://x.org/pdf-guide.html", "https://x.org/get?type=application/pdf", "")],
            "no_lib": P.extract_pdf_text(b"%PDF-1.7 garbage") is None, "sd": sd.to_dict(), "empty": S.extract_structured_data("<html></html>").to_dict()}


# ----------------------------------------------------------------------------- platform
def plugin_registry(pkg, tmp):
    P = _m(pkg, "plugins")
    reg = P.PluginRegistry()
    H = P.HookPoint

    @reg.hook(H.PRE_INDEX)
    def upper(d):
        return {**d, "title": d["title"].upper()}

    def boom(d):
        raise RuntimeError("x")

    def drop(d):
        return None if "spam" in d["title"].lower() else d

    async def bump(x):
        return x + 1

    reg.register_plugin("a", "1.0", {H.PRE_INDEX: boom, H.POST_RANK: bump})
    reg.register_plugin("b", hooks={H.PRE_INDEX: drop, H.POST_RANK: lambda x: x * 10})
    out = {"keep": reg.run_hook(H.PRE_INDEX, {"title": "hello"}), "drop": reg.run_hook(H.PRE_INDEX, {"title": "Spam"}),
           "noop": reg.run_hook(H.POST_SEARCH, [1]), "async": asyncio.run(reg.run_hook_async(H.POST_RANK, 1)),
           "counts": dict(sorted(reg.hook_counts.items())), "plugins": reg.registered_plugins, "points": [p.value for p in H]}
    out["singleton"] = P.get_registry() is P.get_registry()
    return out


def version_tracking(pkg, tmp):
    V = _m(pkg, "version_check")
    parse = {v: V._parse_version(v) for v in ("1.2.3", "1.2.3rc1", "2", "v3.1", "1..4", "garbage", "", "10.0.0-beta.2", "0.1.10", "1.02.3")}
    newer = [V.is_newer(a, b) for a, b in (("1.2.4", "1.2.3"), ("1.2.3", "1.2.3"), ("1.10.0", "1.9.9"), ("0.9", "1.0"), ("1.0.0.1", "1.0.0"), ("2rc1", "2"))]
    t = V.PeerVersionTracker()
    for pid, ver in (("p1", "0.0.1"), ("p2", "999.1.0"), ("p4", ""), ("p5", "999.0.9"), ("p2", "998.0.0")):
        t.record(pid, ver)
    upd = t.check_peer_update()
    with mock.patch.object(V, "_fetch_latest_from_pypi", lambda: None):
        none_yet = V.check_pypi_update(tmp)
    V._write_cache(tmp, "1000.2.3")
    cached = V.check_pypi_update(tmp)
    both = V.check_for_update(tmp, t)
    banner = V.format_update_banner(V.UpdateInfo("1.0", "2.0", "peer")), V.format_update_banner(V.UpdateInfo("1.0", "2.0", "pypi"))
    return {"parse": parse, "newer": newer, "versions": dict(sorted(t.peer_versions.items())), "newest": t.get_newest_peer_version(),
            "upd": (upd.latest, upd.source) if upd else None, "none_yet": none_yet, "cached": (cached.latest, cached.source),
            "both": (both.latest, both.source), "banner": banner, "empty": V.PeerVersionTracker().check_peer_update()}


def dx_helpers(pkg, tmp):
    D = _m(pkg, "dx")

    class Plug:
        name = "demo"

        def __init__(self):
            self.log = []

        def setup(self, app):
            self.log.append(("setup", app))

        def teardown(self):
            self.log.append(("teardown",))

    class Bad:
        name = "bad"

        def setup(self, app):
            raise RuntimeError("nope")

        def teardown(self):
            raise RuntimeError("nope")

    pm, p = D.PluginManager(), Plug()
    pm.register(p, info=D.PluginInfo("demo", "1.0", "a demo"))
    pm.register(Bad())
    pm.setup_all("APP")
    pm.teardown_all()
    infos = [(i.name, i.version, i.description, i.enabled) for i in pm.list_plugins()]
    entry = D.ChangelogEntry("1.2.0", "2030-01-02", ["added x", "fixed y"], ["removed z"])
    return {"log": p.log, "infos": infos, "missing": bool(pm.load_module("definitely.not.a.module")), "no_plugin_var": bool(pm.load_module("json")),
            "tok": [D.DefaultTokenizer().tokenize(s) for s in ("Hello, World", "A big GPU!", "x y zz", "snake_case and kebab-case", "")],
            "guide": D.generate_tool_guide(), "guide_md": D.generate_tool_guide(format="markdown"), "tools": D.MCP_TOOLS_GUIDE,
            "entry": entry.to_markdown(), "plain": D.ChangelogEntry("0.1", "2029-01-01", ["only changes"]).to_markdown(),
            "log_md": D.generate_changelog([entry, D.ChangelogEntry("0.1", "2029-01-01", ["first"])])}


def runtime_files(pkg, tmp):
    import os

    R = _m(pkg, "runtime")
    me = os.getpid()
    ghost = 2 ** 22 + 4321
    out = {"no_pid": R.read_live_pid(tmp), "alive": (R.is_process_running(me), R.is_process_running(ghost), R.is_process_running(0), R.is_process_running(-5))}
    R.write_pid_file(tmp, me)
    out["mine"] = R.read_live_pid(tmp) == me
    R.clear_pid_file(tmp, me + 1)
    out["kept"] = R.pid_path(tmp).exists()
    R.clear_pid_file(tmp, me)
    out["cleared"] = not R.pid_path(tmp).exists()
    R.write_pid_file(tmp, ghost)
    out["stale"] = (R.read_live_pid(tmp), R.pid_path(tmp).exists())
    R.pid_path(tmp).write_text("not a number")
    out["corrupt"] = (R.read_live_pid(tmp), R.pid_path(tmp).exists())
    with R.StartupLock(tmp) as lock:
        second = R.StartupLock(tmp, timeout_seconds=0.05)
        out["lock"] = (lock.acquired, second.acquire(), second.acquired)
    third = R.StartupLock(tmp, timeout_seconds=0.05)
    out["relock"] = third.acquire()
    third.release()
    out["wait_ghost"] = R.wait_for_process_exit(ghost, timeout_seconds=0.05)

    class Level:
        name = "WARNING"

    class Gov:
        degrade_level = Level()
        cpu_percent, memory_percent, process_memory_mb, throttle_factor, checks_performed = 12.345, 45.678, 321.98, 0.87654, 17
        process_memory_limit_mb, process_memory_ratio = 2048, 0.15721

    st = R.build_runtime_status(pid=me, role="full", started_at=1000.0, no_crawl=True, governor_state=Gov())
    out["status"] = {k: v for k, v in st.items() if k not in ("updated_at", "uptime_seconds", "pid")}
    R.write_runtime_status(tmp, dict(st, updated_at=1.0))
    stale = R.read_runtime_status(tmp)
    out["stale_status"] = {k: v for k, v in stale.items() if k not in ("age_seconds", "pid")}
    out["raw_status"] = R.read_runtime_status(tmp, max_age_seconds=None).get("role")
    R.mark_runtime_stopped(tmp, me + 1)
    out["foreign_stop_ignored"] = R.read_runtime_status(tmp, max_age_seconds=None).get("role")
    R.mark_runtime_stopped(tmp, me)
    out["stopped"] = R.read_runtime_status(tmp, max_age_seconds=None).get("status")
    R.runtime_status_path(tmp).write_text("[1, 2]")
    out["not_a_dict"] = (R.read_runtime_status(tmp), R.runtime_status_path(tmp).exists())
    return out


def shutdown_sequence(pkg, tmp):
    S = _m(pkg, "shutdown")
    log = []

    class Ctx:
        async def close_async(self):
            log.append("ctx-async")

        def close(self):
            log.append("ctx-sync")

    class SyncOnly:
        def close(self):
            log.append("ctx-sync")

    async def acb():
        log.append("acb")

    def boom():
        raise RuntimeError("x")

    sd = S.GracefulShutdown()
    sd._context = Ctx()
    for cb in (lambda: log.append("cb1"), boom, acb, lambda: log.append("cb2")):
        sd.add_callback(cb)
    first = (sd.is_shutting_down, sd._try_set_shutting_down(), sd._try_set_shutting_down(), sd.is_shutting_down)
    asyncio.run(sd.cleanup())
    sd2 = S.GracefulShutdown()
    sd2._context = SyncOnly()
    asyncio.run(sd2.cleanup())
    return {"first": first, "log": log}


def persistence_store(pkg, tmp):
    PS = _m(pkg, "persistence.store").PersistentStore
    with PS(tmp / "state.db") as ps:
        for ms in (10.0, 20.0, 40.5):
            ps.record_search(ms)
        ps.record_crawl()
        ps.record_fetch()
        ps.record_fetch()
        ps.register_webhook("https://b.example/hook")
        ps.register_webhook("https://a.example/hook")
        ps.register_webhook("https://b.example/hook")
        ps.save_session("s1", "q1", "r" * 5000)
        ps.save_session("s1", "q2", "short")
        ps.save_session("s2", "q", "x")
        for i in range(5):
            ps.add_history(f"q{i}", i, float(i))
        ps.save_preset("docs", {"language": "en", "limit": 5})
        ps.save_preset("news", {"date_from": 1.0})
        ps.save_preset("docs", {"language": "ko"})
        s1 = ps.get_session("s1")
        out = {"analytics": ps.get_analytics(), "hooks": sorted(ps.get_webhooks()), "unhook": (ps.unregister_webhook("https://a.example/hook"), ps.unregister_webhook("nope")),
               "s1": (s1["last_query"], len(s1["last_results"])), "missing": ps.get_session("zzz"), "expired": (ps.expire_sessions(3600), ps.expire_sessions(-10)),
               "hist": [(h["query"], h["result_count"], h["latency_ms"]) for h in ps.get_history(limit=3)], "cleared": ps.clear_history(),
               "presets": (sorted(ps.list_presets()), ps.get_preset("docs"), ps.get_preset("nope"), ps.delete_preset("news"), ps.delete_preset("news"))}
    with PS(tmp / "state.db") as again:
        out["reopened"] = (again.get_analytics(), again.get_webhooks(), again.list_presets())
    return out


def scalability_helpers(pkg, tmp):
    S = _m(pkg, "scalability")
    bf = S.BloomFilter(capacity=500, fp_rate=0.01)
    items = [f"https://e.example/{i}" for i in range(500)]
    for it in items:
        bf.add(it)
    pool = S.ConnectionPool(str(tmp / "pool.db"), max_connections=1)
    a, b = pool.get(), pool.get()
    a.execute("CREATE TABLE t (x)")
    a.execute("INSERT INTO t VALUES (7)")
    a.commit()
    pool.release(a)
    pool.release(b)          # overflow: closed
    c = pool.get()
    reused, val = c is a, c.execute("SELECT x FROM t").fetchone()[0]
    pool.release(c)
    try:
        b.execute("SELECT 1")
        overflow_closed = False
    except Exception:  # noqa: BLE001
        overflow_closed = True
    pool.close_all()

    class Store:
        def __init__(self):
            self.rows = []

        def add_document(self, **kw):
            if not kw["url"].startswith("https://"):
                raise ValueError("bad scheme")
            self.rows.append(kw)

    st = Store()
    res = S.batch_ingest(st, [{"url": "https://a.example/1", "title": "T", "text": "alpha"}, {"url": "https://a.example/2", "content": "beta", "text_hash": "given"},
                              {"url": "ftp://nope", "text": "x"}, {"title": "no url"}, {"url": "https://a.example/3", "text": "gamma", "content_hash": "rawh", "language": "ko"}])
    return {"no_false_negatives": all(it in bf for it in items), "len": len(bf), "size_bytes": bf.size_bytes, "fp_bounded": sum(f"https://o.example/{i}" in bf for i in range(2000)) < 80,
            "pool": (reused, val, overflow_closed), "ingest": (res.total, res.succeeded, res.failed, [e.split(":")[0] for e in res.errors]),
            # (hashes missing from a record default to the text digest here and to "" in the reference, which then folds the batch into one
            #  row through the store's uniqueness check: deliberately different, so only explicitly given hashes are compared)
            "rows": [(r["url"], r["title"], r["text"], r["language"]) for r in st.rows], "given_hash": st.rows[1]["text_hash"], "given_raw": st.rows[2]["raw_html_hash"]}


def security_operations(pkg, tmp):
    SO = _m(pkg, "security_ops")
    f = tmp / "keys" / "api_keys.json"
    m = SO.APIKeyManager(f)
    m.add_key("alpha-secret", "alpha")
    m.add_key("temp-secret", ttl_days=1)
    checks = [m.validate(k) for k in ("alpha-secret", "temp-secret", "nope", "")]
    rotated = m.add_key("beta-secret", "alpha-rotated-1")
    after_rotate = [m.validate(k) for k in ("alpha-secret", "beta-secret")]
    revoked = (m.revoke("key-2"), m.revoke("ghost"), m.validate("temp-secret"))
    listing = [(e["label"].split("-rotated-")[0], e["revoked"], e["active"], e["expires"] is None) for e in m.list_keys()]
    again = SO.APIKeyManager(f)
    persisted = ([again.validate(k) for k in ("alpha-secret", "beta-secret", "temp-secret")], len(again.list_keys()), "secret" in f.read_text())
    f.write_text("{broken")
    a = SO.AuditLogger(tmp / "audit" / "audit.log", max_size_mb=1)
    a.log("search", details="q=" + "x" * 900)
    a.log("crawl", source="mcp", client="10.0.0.1", success=False)
    # (a corrupt line ends the reference's read-back; this repo skips it and keeps going: not compared)
    a.log("fetch", source="cli")
    rec = a.recent(limit=10)
    return {"checks": checks, "rotated_label": rotated.label.split("-rotated-")[0], "after_rotate": after_rotate, "revoked": revoked, "listing": listing,
            "persisted": persisted, "corrupt": SO.APIKeyManager(f).list_keys(), "audit": [(e.action, e.source, e.client, len(e.details), e.success) for e in rec],
            "limit1": [e.action for e in a.recent(limit=1)], "nolog": SO.AuditLogger(None).recent(), "nofile": SO.AuditLogger(tmp / "never.log").recent()}


# ----------------------------------------------------------------------------- p2p / trust
def peer_profiles(pkg, tmp):
    PP = _m(pkg, "p2p.peer_profile")
    t = PP.PeerProfileTracker()
    samples = {"fast": [20, 30, 25, 22, 40, 18], "mid": [150, 180, 140, 400, 120], "slow": [900, 800, 1000, 2500], "new": [10]}
    for pid, xs in samples.items():
        for i, ms in enumerate(xs):
            t.record(pid, float(ms), success=(i % 4 != 3))
    snap = {pid: (round(p.avg_latency_ms, 6), round(p.p95_latency_ms, 6), round(p.success_rate, 6), p.bandwidth_class.value, p.total_interactions)
            for pid in samples if (p := t.get(pid))}
    with mock.patch("random.random", return_value=0.99):
        calm = t.rank_by_latency(["slow", "ghost", "fast", "mid", "new"])
    with mock.patch("random.random", return_value=0.0):
        shaken = t.rank_by_latency(["slow", "ghost", "fast", "mid", "new"])
    return {"snap": snap, "plain": t.rank_by_latency(["slow", "ghost", "fast", "mid", "new"], diversity=False), "calm": calm, "shaken": shaken,
            "short": t.rank_by_latency(["slow", "fast"]), "timeouts": {p: t.adaptive_timeout(p) for p in ("fast", "mid", "slow", "ghost")},
            "timeouts5k": {p: t.adaptive_timeout(p, base_ms=5000) for p in ("fast", "mid", "slow")},
            "pct": [PP._percentile(v, q) for v, q in (([1, 2, 3, 4], 50), ([], 95), ([5], 95), ([1, 2, 3, 4, 5, 6, 7, 8, 9, 10], 95), ([3, 1, 2], 0), ([3, 1, 2], 100))],
            "cls": [PP._classify_bandwidth(x).value for x in (0, 99.9, 100, 499.9, 500, 1e6)], "known": t.known_peers, "default": t.get_or_default("zzz").bandwidth_class.value}


def load_guard(pkg, tmp):
    LG = _m(pkg, "p2p.load_guard")
    g = LG.NodeLoadGuard(max_queries_per_minute=4, max_concurrent=2)
    trace = [g.try_acquire("a"), g.try_acquire("b"), g.try_acquire("c")]
    g.release("a")
    trace += [g.try_acquire("a"), g.is_overloaded]
    g.release("a")
    g.release("b")
    trace += [g.try_acquire("c"), g.try_acquire("c")]
    st = g.stats
    info = g.get_reject_info()
    out = {"trace": trace, "stats": (st.accepted, st.rejected, st.concurrent, st.queries_this_minute, st.is_overloaded), "info": info,
           "per_peer": {p: g.peer_query_count(p) for p in "abcz"}}
    for _ in range(5):
        g.release()
    g.reset()
    st = g.stats
    out["reset"] = (st.accepted, st.rejected, st.concurrent, st.queries_this_minute, st.is_overloaded, g.peer_query_count("a"), g.try_acquire())
    return out


def merkle_tree(pkg, tmp):
    import hashlib

    M = _m(pkg, "trust.merkle")
    out = {}
    for n in (1, 2, 3, 7, 16):
        hs = [hashlib.sha256(f"doc{i}".encode()).hexdigest() for i in range(n)]
        t = M.MerkleTree()
        root = t.build(hs)
        proofs = [M.serialize_proof(t.get_proof(i)) for i in (0, n - 1, n // 2)]
        out[n] = {"root": root, "height": t.height, "leaves": t.leaf_count, "paths": [[(h, str(s)) for h, s in p["proof_path"]] for p in proofs],
                  "ok": [M.MerkleTree.verify_document(hs[i], t.get_proof(i)) for i in range(n)],
                  "tampered": M.MerkleTree.verify_document("f" * 64, t.get_proof(0))}
    t = M.MerkleTree()
    errs = []
    for fn in (lambda: t.build([]), lambda: t.get_proof(0)):
        try:
            fn()
            errs.append(None)
        except Exception as e:  # noqa: BLE001
            errs.append(type(e).__name__)
    t.build(["a" * 64])
    try:
        t.get_proof(5)
    except Exception as e:  # noqa: BLE001
        errs.append(type(e).__name__)
    rec = t.create_root_record("peer-x")
    wire = M.serialize_merkle_root(rec)
    back = M.deserialize_merkle_root(wire)
    out["errors"] = errs
    out["root_record"] = ({k: v for k, v in wire.items() if k != "built_at"}, back.root_hash == rec.root_hash, back.signature)
    return out


def threat_detector(pkg, tmp):
    D = _m(pkg, "trust.detector")
    TS = _m(pkg, "trust.scoring")
    FD = _m(pkg, "credits.farming")

    class PT:
        def __init__(self, score, tier, fails, isolated=False):
            self.trust_score, self.tier, self.consecutive_audit_failures, self.isolated = score, tier, fails, isolated

    class Trust:
        def __init__(self):
            self.isolated = []
            self.table = {"good": PT(0.9, TS.TrustTier.TRUSTED, 0), "weak1": PT(0.45, TS.TrustTier.NORMAL, 0), "weak2": PT(0.45, TS.TrustTier.NORMAL, 3),
                          "untrusted": PT(0.1, TS.TrustTier.UNTRUSTED, 0), "iso": PT(0.0, TS.TrustTier.UNTRUSTED, 9, True), "farm": PT(0.8, TS.TrustTier.NORMAL, 0),
                          "anom": PT(0.8, TS.TrustTier.NORMAL, 2), "rate": PT(0.8, TS.TrustTier.NORMAL, 0)}

        def get_trust(self, pid):
            return self.table.get(pid)

        def isolate_peer(self, pid):
            self.isolated.append(pid)

    class FC:
        def __init__(self, verdict, anomalies=0, rl=False):
            self.verdict, self.anomaly_count, self.rate_limit_exceeded = verdict, anomalies, rl

    class Farm:
        def check(self, pid, action):
            V = FD.FarmingVerdict
            return {"farm": FC(V.BLOCKED), "anom": FC(V.SUSPICIOUS, 2), "rate": FC(V.RATE_LIMITED, 0, True)}.get(pid, FC(V.CLEAN))

    trust = Trust()
    det = D.MaliciousNodeDetector(trust, Farm())
    res = {}
    for pid in ("good", "weak1", "weak2", "untrusted", "iso", "farm", "anom", "rate", "stranger"):
        a = det.assess_and_enforce(pid)
        res[pid] = (a.threat_level.value, a.should_isolate, a.weak_signals, a.detail, a.trust_score, a.trust_tier.value, a.farming_verdict.value,
                    a.consecutive_audit_failures, a.anomaly_count)
    return {"res": res, "isolated": trust.isolated}


def ranking_and_remote(pkg, tmp):
    R = _m(pkg, "index.ranking")
    Q = _m(pkg, "search.query")
    now = 1_700_000_000.0
    cands = [R._RawCandidate(doc_id=i, url=f"https://e.example/{i}", title=f"T{i}", snippet="s", bm25_raw=b, crawled_at=now - age, peer_id=None, trust=t, authority=a,
                             title_match=tm, url_path=up)
             for i, (b, age, t, a, tm, up) in enumerate([(5.0, 0, 0.5, 0.0, 0.0, 0.0), (4.0, 86400 * 7, 0.9, 0.5, 1.0, 0.0), (0.0, 0, 0.5, 0.0, 0.0, 0.0),
                                                         (9.0, 86400 * 365, 0.1, 0.0, 0.0, 1.0), (5.0, 0, 0.5, 0.0, 0.0, 0.0), (2.5, -50, 0.5, 1.0, 0.5, 0.5)])]
    ranked = R.rank_results(cands, limit=4, now=now)
    weighted = R.rank_results(list(reversed(cands)), limit=6, now=now)
    weird = [None, True, "12", "1e3", "nan", "inf", float("inf"), -3.9, [1], {"a": 1}, "abc", 2 ** 70, b"7"]
    return {"ranked": [(r.doc_id, r.bm25_score, r.freshness_score, r.trust_score, r.authority_score, r.title_match_score, r.url_path_score, r.combined_score) for r in ranked],
            "weighted": [r.doc_id for r in weighted], "empty": R.rank_results([], now=now),
            "ints": [Q._safe_remote_int(w) for w in weird], "ints_d": Q._safe_remote_int("x", default=-1),
            "floats": [Q._safe_remote_float(w) for w in weird],
            "sanitize": [Q._sanitize_fts_query(q) for q in ('"quoted phrase" AND (x OR y)', "near NEAR far", "***((()))", "   ", "c++ && c#", "日本語 検索", "a" * 1200 + " tail",
                                                             "NOT", "android", "ORacle NOTe", "{x:y}^2", "tab\tnew\nline")],
            "score": R.combined_score(0.5, 0.5, 0.5, 0.5, title_match=1.0, url_path=1.0), "fresh": [R.freshness_score(now - a, now=now) for a in (0, 604800, 6048000, -5)],
            "norm": [R.normalize_bm25(s, max_score=m) for s, m in ((3.0, 3.0), (0.0, 3.0), (-1.0, 3.0), (3.0, 0.0))]}


def timezone_table(pkg, tmp):
    T = _m(pkg, "credits.timezone_verify")
    ips = ["1.2.3.4", "8.8.8.8", "41.0.0.1", "58.1.1.1", "61.2.3.4", "77.1.2.3", "103.5.6.7", "133.9.9.9", "150.1.1.1", "175.2.2.2", "185.3.3.3", "190.4.4.4",
           "193.5.5.5", "200.6.6.6", "202.7.7.7", "210.8.8.8", "218.9.9.9", "223.1.1.1", "10.0.0.1", "not-an-ip", "", "256.1.1.1", "::1"]
    fn = getattr(T, "estimate_offset_from_ip", None) or getattr(T, "ip_to_utc_offset", None)
    out = {"offsets": {tz: T.get_timezone_offset(tz) for tz in ("Asia/Seoul", "UTC", "Europe/Berlin", "America/Los_Angeles", "Asia/Kolkata", "Nope/Zone", "")}}
    if fn is not None:
        out["ips"] = {ip: fn(ip) for ip in ips}
    return out


SCENARIOS = {f.__name__: f for f in (
    scheduler_bookkeeping, freshness_queue, feed_monitor, crawl_intelligence, url_assignment, pdf_and_structured, plugin_registry, version_tracking,
    dx_helpers, runtime_files, shutdown_sequence, persistence_store, scalability_helpers, security_operations, peer_profiles, load_guard, merkle_tree,
    threat_detector, ranking_and_remote, timezone_table)}


# ----------------------------------------------------------------------------- second batch: stores and accounting
def query_cache(pkg, tmp):
    C = _m(pkg, "search.cache")
    c = C.QueryCache(max_size=3, ttl_seconds=1000.0)
    trace = [c.get("a", 10)]
    c.put("a", 10, [1, 2])
    c.put("A ", 10, [9])                 # normalisation: same key or not -- both sides must agree
    c.put("b", 10, [3])
    c.put("b", 5, [4])
    trace += [c.get("a", 10), c.get("a", 5), c.get("b", 10), c.get("b", 5), c.size]
    c.put("c", 10, [5])
    c.put("d", 10, [6])                  # evicts the least recently used
    trace += [c.get("a", 10), c.get("b", 10), c.get("c", 10), c.get("d", 10), c.size, c.invalidate("c", 10), c.invalidate("zzz", 1), c.size]
    st = c.stats
    out = {"trace": trace, "stats": (st.hits, st.misses, st.evictions, st.total, round(st.hit_rate, 6))}
    old = C.QueryCache(max_size=5, ttl_seconds=-1.0)
    old.put("x", 1, [1])
    out["expired"] = (old.get("x", 1), old.evict_expired(), old.size)
    c.clear()
    out["cleared"] = c.size
    return out


def credit_ledger(pkg, tmp):
    L = _m(pkg, "credits.ledger")
    T = _m(pkg, "credits.types")
    led = L.CreditLedger(tmp / "credits.db", owner_email="dev@example.org")
    A = T.ActionType
    earned = [led.record_action(A.CRAWL, 10, note="ten pages"), led.record_action(A.QUERY_PROCESS, 4), led.record_action(A.LLM_SUMMARIZE_OWN, 2, off_peak=True),
              led.record_action(A.DOC_HOSTING, 50), led.record_action(A.NETWORK_UPTIME, 3.5), led.record_action(A.LLM_SUMMARIZE_PEER, 100)]
    spends = [led.spend(1.0), led.spend(0.25, reason="search"), led.spend(10_000.0)]
    st = led.stats()
    al = led.search_allowance()
    out = {"earned": earned, "spends": spends, "totals": (round(led.total_earned(), 6), round(led.total_spent(), 6), round(led.balance(), 6), round(led.debt_amount(), 6)),
           "score": round(led.contribution_score(), 6), "tier": led.tier().value, "cost": led.search_cost(), "state": led.credit_state().value,
           "allowance": (al.state.value, al.search_cost, al.debt_amount), "stats": (round(st.total_earned, 6), round(st.total_spent, 6), round(st.balance, 6)),
           "by_action": sorted((a, round(v, 6)) for a, v in led.earnings_by_action()),
           "entries": [(e.action, e.quantity, e.weight, e.multiplier, round(e.credits, 6), e.note) for e in led.recent_entries(limit=3)],
           "owner": led.owner_email, "off_peak": [L.is_off_peak(hour=h) for h in (0, 6, 7, 12, 22, 23)]}
    led.close()
    poor = L.CreditLedger(tmp / "poor.db")
    out["poor"] = (poor.balance(), poor.spend(0.5), poor.credit_state().value, poor.search_allowance().state.value, poor.tier().value, poor.search_cost())
    poor.close()
    return out


def trust_store(pkg, tmp):
    S = _m(pkg, "trust.scoring")
    ts = S.TrustStore(tmp / "trust.db")
    ts.update_uptime("good", 200.0)
    ts.update_contribution("good", 500.0)
    for ok in (True, True, True, False, True):
        ts.record_audit("good", passed=ok)
    ts.record_summary_rating("good", 0.9)
    for _ in range(4):
        ts.record_audit("bad", passed=False)
    ts.update_uptime("idle", 1.0)
    ts.isolate_peer("bad")
    def view(pid):
        t = ts.get_trust(pid)
        return None if t is None else (round(t.uptime_score, 6), round(t.contribution_score, 6), round(t.audit_pass_rate, 6), round(t.summary_quality, 6),
                                       round(t.trust_score, 6), t.tier.value, t.consecutive_audit_failures, t.isolated)
    out = {"good": view("good"), "bad": view("bad"), "idle": view("idle"), "none": view("ghost"), "score_ghost": ts.get_trust_score("ghost"),
           "listed": sorted(t.peer_id for t in ts.list_peers()), "listed_all": sorted(t.peer_id for t in ts.list_peers(include_isolated=True)),
           "isolated": [t.peer_id for t in ts.list_isolated()], "is_iso": (ts.is_isolated("bad"), ts.is_isolated("good"))}
    ts.unisolate("bad")
    out["after_unisolate"] = (ts.is_isolated("bad"), view("bad"))
    out["formula"] = [round(S.compute_trust_score(*a), 6) for a in ((0, 0, 0, 0, 0.0), (100, 50, 10, 9, 0.8), (10_000, 10_000, 100, 100, 1.0), (24, 3, 4, 1, 0.2), (720, 1000, 0, 0, 0.5))]
    out["tiers"] = [S.trust_tier(x).value for x in (0.0, 0.29, 0.3, 0.59, 0.6, 0.79, 0.8, 1.0)]
    return out


def dedup_store(pkg, tmp):
    D = _m(pkg, "crawler.dedup")
    H = _m(pkg, "hashing")
    db = D.DeduplicatorDB(str(tmp / "dedup.db"))
    text = "Tensor memory holds the accumulators of the fifth generation tensor cores. " * 6
    near = text.replace("fifth", "5th", 1)
    other = "Completely unrelated cooking instructions about pasta and tomato sauce for dinner tonight. " * 6
    trace = [db.is_url_seen("https://Example.org/a?utm_source=x"), db.is_content_seen(H.content_hash(text)), db.is_near_duplicate(text)]
    db.mark_seen("https://example.org/a", H.content_hash(text), text)
    trace += [db.is_url_seen("https://Example.org/a?utm_source=x"), db.is_url_seen("https://example.org/a/"), db.is_url_seen("https://example.org/b"),
              db.is_content_seen(H.content_hash(text)), db.is_content_seen(H.content_hash(near)), db.is_near_duplicate(near), db.is_near_duplicate(other),
              db.is_near_duplicate(near, threshold=0)]
    db.mark_seen("https://example.org/b", H.content_hash(other), other, commit=False)
    db.flush()
    db.close()
    again = D.DeduplicatorDB(str(tmp / "dedup.db"))
    trace += [again.is_url_seen("https://example.org/b"), again.is_near_duplicate(near), again.is_content_seen("nope")]
    again.close()
    return {"trace": trace}


def peer_store(pkg, tmp):
    P = _m(pkg, "p2p.peer_store")
    ps = P.PeerStore(tmp)
    ps.upsert("p1", "/ip4/10.0.0.1/tcp/4001")
    ps.upsert("p2", "/ip4/10.0.0.2/tcp/4001")
    ps.upsert("p1", "/ip4/10.0.0.9/tcp/4001")
    ps.record_failure("p2")
    ps.record_failure("p2")
    ps.record_failure("ghost")
    ps.save_connected([("p3", "/ip4/10.0.0.3/tcp/4001"), ("p1", "/ip4/10.0.0.9/tcp/4001")])
    rows = {c.peer_id: (c.multiaddr, c.success_count, c.fail_count, round(c.success_rate, 6)) for c in ps.load_recent(limit=10)}
    out = {"rows": rows, "count": ps.count(), "limit1": len(ps.load_recent(limit=1))}
    ps.remove("p3")
    out["after_remove"] = (ps.count(), sorted(c.peer_id for c in ps.load_recent()))
    out["prune_none"] = ps.prune(max_age_hours=1000, max_peers=10)
    out["prune_cap"] = (ps.prune(max_age_hours=1000, max_peers=1), ps.count())
    out["prune_age"] = (ps.prune(max_age_hours=-1.0), ps.count())
    return out


def feedback_store(pkg, tmp):
    F = _m(pkg, "search.feedback")
    fs = F.FeedbackStore(str(tmp / "fb.db"))
    fs.record_fetch("gpu tensor cores", "https://a.example/1", 1)
    fs.record_fetch("gpu tensor cores", "https://a.example/1", 3)
    fs.record_fetch("other", "https://a.example/2", 9)
    fs.record_skip("gpu tensor cores", ["https://a.example/2", "https://a.example/3"])
    fs.record_citation("gpu tensor cores", "https://a.example/1")
    fs.record_reformulation("gpu tensor cores")
    def boost(u):
        b = fs.get_url_stats(u)
        return None if b is None else (round(b.boost_score, 6), b.fetch_count, b.skip_count, b.cite_count)
    out = {"boosts": {u: (round(fs.get_boost(u), 6), boost(u)) for u in ("https://a.example/1", "https://a.example/2", "https://a.example/3", "https://nobody.example/")},
           "signals": fs.signal_count(), "top": [(b.url, round(b.boost_score, 6)) for b in fs.top_boosted_urls(limit=2)],
           "reform": (fs.is_reformulation("gpu tensor cores"), fs.is_reformulation("never asked"), fs.is_reformulation("gpu tensor cores", window=-1.0))}
    fs.close()
    return out


SCENARIOS.update({f.__name__: f for f in (query_cache, credit_ledger, trust_store, dedup_store, peer_store, feedback_store)})


# ----------------------------------------------------------------------------- third batch: detectors, graph, exchange, metrics
def farming_detector(pkg, tmp):
    F = _m(pkg, "credits.farming")
    fd = F.FarmingDetector(tmp / "farm.db")
    t0 = 1_000_000.0
    fd.register_node("new", now=t0)
    fd.register_node("old", now=t0 - 10 * 86400)
    out = {"probation": (fd.is_on_probation("new", now=t0 + 3600), fd.is_on_probation("old", now=t0), fd.is_on_probation("ghost", now=t0),
                         round(fd.probation_remaining("new", now=t0 + 3600), 4), round(fd.probation_remaining("old", now=t0), 4))}
    for i in range(30):                       # perfectly regular crawl actions, one every 60 s
        fd.log_action("bot", "crawl", now=t0 + i * 60.0)
    rng = [3, 41, 97, 160, 171, 305, 420, 455, 610, 777, 905, 1111]
    for off in rng:                           # irregular human-ish pattern
        fd.log_action("human", "crawl", now=t0 + off)
    for i in range(200):                      # burst: 200 actions inside one minute
        fd.log_action("burst", "crawl", now=t0 + 1700 + i * 0.2)
    now = t0 + 1800
    out["counts"] = {p: fd.actions_in_last_hour(p, "crawl", now=now) for p in ("bot", "human", "burst", "ghost")}
    out["regular"] = {p: fd.detect_regular_intervals(p, "crawl", now=now) for p in ("bot", "human", "burst")}
    out["bursts"] = {p: fd.detect_burst(p, "crawl", now=now) for p in ("bot", "human", "burst")}
    out["rate_limited"] = {p: fd.is_rate_limited(p, "crawl", now=now) for p in ("bot", "human", "burst")}
    checks = {}
    for p in ("old", "new", "bot", "human", "burst"):
        if p in ("bot", "human", "burst"):
            fd.register_node(p, now=t0 - 30 * 86400)
        c = fd.check(p, "crawl", now=now)
        checks[p] = (c.verdict.value, c.rate_limit_exceeded, c.anomaly_count, round(c.probation_remaining_hours, 3))
    out["checks"] = checks
    ids = [fd.record_anomaly("bad", "manual", f"n{i}", now=now + i) for i in range(4)]
    hist = fd.get_anomaly_history("bad", limit=2)
    out["anomalies"] = (len(ids), ids == sorted(ids), [(a.anomaly_type, a.detail) for a in hist], fd.is_blocked("bad"), fd.check("bad", "crawl", now=now + 10).verdict.value)
    fd.unblock("bad")
    out["unblocked"] = fd.is_blocked("bad")
    out["pruned"] = fd.prune_old_actions(max_age_seconds=0.0) > 0
    return out


def link_graph(pkg, tmp):
    G = _m(pkg, "index.link_graph")
    g = G.LinkGraph(str(tmp / "links.db"))
    added = [g.add_links("https://a.example/1", ["https://b.example/x", "https://c.example/y", "https://a.example/2", "https://b.example/x"]),
             g.add_links("https://b.example/x", ["https://c.example/y", "https://d.example/"]),
             g.add_links("https://c.example/y", ["https://a.example/1"]),
             g.add_links("https://e.example/", ["https://c.example/z", "not a url", ""]),
             g.add_links("https://a.example/1", ["https://b.example/x"])]
    auth = g.compute_domain_authority()
    # (`add_links` returns the number of NEW edges here and the number of valid targets in the reference: not compared)
    out = {"n_calls": len(added), "stats": g.get_stats(), "domains": sorted(auth), "sum": round(sum(auth.values()), 6),
           "order": [d for d, _ in sorted(auth.items(), key=lambda kv: (-round(kv[1], 9), kv[0]))],
           "values": {d: round(v, 6) for d, v in auth.items()},
           "lookups": (round(g.domain_authority("c.example"), 6), g.domain_authority("nobody.example"), round(g.url_authority("https://c.example/anything"), 6), g.url_authority("garbage"))}
    g.close()
    return out


def peer_exchange(pkg, tmp):
    X = _m(pkg, "p2p.pex")
    ex = X.PeerExchange("me")
    peers = [(f"p{i}", f"/ip4/10.0.0.{i}/tcp/4001") for i in range(1, 16)] + [("me", "/ip4/10.0.0.99/tcp/4001")]
    resp = ex.build_response(peers, max_peers=5)
    full = ex.build_response(peers[:3])
    got = ex.process_response("sender", [{"peer_id": "n1", "multiaddr": "/ip4/1.1.1.1/tcp/4001"}, {"peer_id": "me", "multiaddr": "/ip4/2.2.2.2/tcp/1"},
                                           {"peer_id": "known", "multiaddr": "/ip4/3.3.3.3/tcp/1"}, {"peer_id": "", "multiaddr": "/ip4/4.4.4.4/tcp/1"},
                                           {"peer_id": "n2"}, {"peer_id": "n3", "multiaddr": 42}, {"peer_id": "n1", "multiaddr": "/ip4/1.1.1.1/tcp/4001"},
                                           {"peer_id": "sender", "multiaddr": "/ip4/5.5.5.5/tcp/1"}], known_peers={"known"})
    # (a non-dict entry in peers_data makes the reference raise AttributeError; this repo skips it: not compared)
    limits = [ex.check_rate_limit("asker") for _ in range(8)]
    ex.cleanup_rate_limits()
    return {"resp_len": len(resp), "resp_keys": sorted(resp[0]) if resp else [], "no_self": all(r["peer_id"] != "me" for r in resp), "full": full,
            "got": [(p.peer_id, p.multiaddr) for p in got], "limits": limits, "other": ex.check_rate_limit("someone else")}


def metrics_and_slo(pkg, tmp):
    M = _m(pkg, "observability.metrics")
    S = _m(pkg, "slo")
    mc = M.MetricsCollector()
    mc.inc("searches_total")
    mc.inc("searches_total", 2)
    mc.set_gauge("peers_connected", 7)
    for v in (5.0, 15.0, 250.0, 1200.0):
        mc.observe("search_latency_ms", v)
    d = mc.to_dict()
    prom = mc.format_prometheus()
    tr = M.QueryTrace("t1", "gpu")
    tr.add_span(M.QuerySpan("s1", "peerA", "local", 1.0, 1.5, 500.0, {"k": "v"}))
    tr.add_span(M.QuerySpan("s2", "peerB", "remote", 1.1, 1.9, 800.0))
    rules = M.generate_alert_rules()
    dash = M.generate_grafana_dashboard()
    t = S.SLOTracker([S.SLODefinition("lat", "p95 latency", 100.0, "ms", 3600.0), S.SLODefinition("avail", "availability", 0.99, "ratio", 3600.0)])
    for v in (50.0, 80.0, 400.0):
        t.record("lat", v)
    for ok in (True, True, True, False):
        t.record_success("avail", ok)
    t.record("unknown", 1.0)
    st = {s.slo.name: (round(s.current_value, 6), s.target, s.met, round(s.error_budget_remaining, 6)) for s in t.get_status()}
    summ = t.summary()
    d.pop("uptime_seconds", None)
    plain = sorted(ln for ln in prom.splitlines() if ln and not ln.startswith("#") and "quantile" not in ln)     # quantile series are an addition here
    fwd = M.configure_log_forwarding(format="json", output="stdout")
    return {"dict": d, "prom_lines": plain, "trace": tr.to_dict(), "rules": [(r.get("alert"), r.get("expr")) for r in rules][:5],      # GPU alert rules are additions
            "dash_keys": sorted(dash)[:3], "slo": st, "slo_summary_keys": sorted(summ), "log_fwd": {k: fwd[k] for k in ("format", "output", "level")}}


def sessions_and_webhooks(pkg, tmp):
    S = _m(pkg, "mcp.session")
    a = S.AnalyticsTracker()
    for ms in (10.0, 30.0, 20.0):
        a.record_search(ms)
    a.record_crawl()
    a.record_fetch()
    a.record_fetch()
    store = S.SessionStore(max_size=2, ttl_seconds=1000.0)
    s1 = store.get_or_create("s1")
    fields = sorted(k for k in dir(s1) if not k.startswith("_") and k in ("last_query", "last_results", "updated_at"))
    same = kept = None       # (object identity across calls differs between the implementations' stores: not compared)
    wh = S.WebhookRegistry(max_registrations=2)
    regs = [wh.register("https://hooks.example/a"), wh.register("https://hooks.example/a"), wh.register("http://127.0.0.1/x"), wh.register("ftp://hooks.example/b"),
            wh.register("https://hooks.example/b"), wh.register("https://hooks.example/c")]
    return {"analytics": a.to_dict(), "session_fields": fields, "regs": [r is None for r in regs], "urls": sorted(wh.urls),
            "unreg": (wh.unregister("https://hooks.example/a"), wh.unregister("https://hooks.example/zzz")), "left": sorted(wh.urls)}


SCENARIOS.update({f.__name__: f for f in (farming_detector, link_graph, peer_exchange, metrics_and_slo, sessions_and_webhooks)})


# ----------------------------------------------------------------------------- fourth batch: recrawl, sybil limits, attestations, scheduling, snapshots
def recrawl_selection(pkg, tmp):
    R = _m(pkg, "crawler.recrawl")
    now = 1_000_000.0

    def cand(i, **kw):
        base = dict(doc_id=i, url=f"https://e.example/{i}", text_hash="h", etag=None, last_modified=None, recrawl_interval=1000, stale_count=0,
                    change_frequency=0.2, crawled_at=now - 5000.0, last_recrawl_at=None)
        base.update(kw)
        return R.RecrawlCandidate(**base)

    docs = [cand(1), cand(2, last_recrawl_at=now - 100), cand(3, last_recrawl_at=now - 50_000), cand(5, recrawl_interval=10_000_000),
            cand(6, crawled_at=now - 999), cand(7, crawled_at=now - 1001), cand(8, last_recrawl_at=now - 2000, recrawl_interval=500), cand(9, stale_count=2, crawled_at=now - 9e6)]
    # (a candidate that already reached the stale threshold is skipped here and still selected by the reference, whose store
    #  query filters it earlier: not compared)
    return {"picked": [c.doc_id for c in R.select_candidates(docs, now=now)], "top2": [c.doc_id for c in R.select_candidates(docs, now=now, max_batch=2)],
            "none": R.select_candidates([], now=now), "intervals": [R.compute_recrawl_interval(x / 20) for x in range(0, 21)],
            "freq": [round(R.update_change_frequency(f, ch, alpha=a), 6) for f, ch, a in ((0.5, True, 0.3), (0.5, False, 0.3), (0.0, True, 0.3), (1.0, False, 0.3), (0.2, True, 1.0), (0.9, True, 0.0))]}


def subnet_limits(pkg, tmp):
    S = _m(pkg, "p2p.sybil")
    lim = S.SubnetLimiter(max_per_subnet=2)
    trace = [lim.add("10.1.2.3", "a", 0), lim.add("10.1.2.4", "b", 0), lim.can_add("10.1.2.5", 0), lim.add("10.1.2.5", "c", 0), lim.add("10.1.3.5", "c", 0),
             lim.add("10.1.2.9", "d", 1), lim.add("2001:db8::1", "v6", 0), lim.add("2001:db8::2", "v6b", 0)]
    # (re-adding a peer that is already counted in a full subnet is accepted here, idempotently, and refused by the reference;
    #  an unparsable address raises ValueError in the reference and is refused here: neither is compared)
    counts = {b: dict(sorted(lim.get_subnet_counts(b).items())) for b in (0, 1, 2)}
    lim.remove("10.1.2.3", "a", 0)
    lim.remove("10.9.9.9", "ghost", 0)
    after = (lim.can_add("10.1.2.5", 0), lim.total_nodes())
    key = bytes(range(32))
    pow_ = S.generate_pow(key, difficulty_bits=8)
    v = S.SybilValidator(difficulty_bits=8, max_per_subnet=1)
    checks = [v.validate_peer(key, pow_.nonce, "192.0.2.1", "p1", 3), v.validate_peer(key, pow_.nonce, "192.0.2.7", "p2", 3), v.validate_peer(key, pow_.nonce + 1, "198.51.100.1", "p3", 3)]
    return {"trace": trace, "counts": counts, "after": after, "pow_ok": (S.verify_pow(key, pow_.nonce, 8), pow_.difficulty_bits, pow_.hash_hex == S.compute_pow_hash(key, pow_.nonce).hex()),
            "node_id": S.derive_node_id(key, pow_.nonce), "checks": [(ok, why.split(":")[0][:40]) for ok, why in checks[:2]] + [checks[2][0] in (True, False)]}


def attestations(pkg, tmp):
    A = _m(pkg, "trust.attestation")
    K = _m(pkg, "p2p.keys")
    kp = K.KeyPair.generate()
    raw, text = b"<html><body>Tensor memory is 256 KB per SM.</body></html>", "Tensor memory is 256 KB per SM."
    att = A.create_attestation("https://e.example/a", raw, text, kp, crawled_at=1_700_000_000.0)
    wire = A.serialize_attestation(att)
    back = A.deserialize_attestation(wire)
    res = {name: A.verify_attestation(back, kp, **kw) for name, kw in (("both", dict(raw_body=raw, extracted_text=text)), ("raw_only", dict(raw_body=raw)),
                                                                        ("tampered_text", dict(raw_body=raw, extracted_text=text + "!")), ("tampered_raw", dict(raw_body=raw + b" ")),
                                                                        ("nothing", {}))}
    other = K.KeyPair.generate()
    forged = A.verify_attestation(back, other, raw_body=raw, extracted_text=text)
    return {"fields": (att.url, att.raw_hash, att.text_hash, att.crawled_at, att.content_length, len(att.signature)), "wire_keys": sorted(wire), "round_trip": back == att,
            "results": {k: (r.raw_match, r.text_match, r.signature_valid, r.verified) for k, r in res.items()},
            "forged": (forged.signature_valid, forged.verified)}


def llm_scheduling(pkg, tmp):
    S = _m(pkg, "credits.scheduling")
    N = S.NodeScheduleInfo
    nodes = [N("day", 23, 7, "UTC", True, 0.9), N("night", 10, 18, "UTC", True, 0.6), N("nollm", 10, 18, "UTC", False, 1.0), N("wrap", 22, 6, "UTC", True, 0.7),
             N("low", 10, 18, "UTC", True, 0.2)]
    sch = S.EnergyAwareScheduler()
    out = {}
    for hour in (3, 12, 20, 23):
        d = sch.schedule_llm_task(nodes, now_override_hour=hour)
        out[hour] = None if d is None else (d.target_peer_id, d.is_off_peak, d.credit_multiplier)
    batch = sch.schedule_batch(nodes, 5, now_override_hour=12)
    return {"single": out, "batch": [(d.target_peer_id, d.is_off_peak, d.credit_multiplier) for d in batch], "empty": sch.schedule_llm_task([], now_override_hour=1),
            "no_llm": sch.schedule_llm_task([nodes[2]], now_override_hour=12), "off_peak": [S.node_is_off_peak(n, now_override_hour=12) for n in nodes],
            "window": [S.is_off_peak_at(hour=h, start=22, end=6) for h in (21, 22, 0, 5, 6, 12)], "same": [S.is_off_peak_at(hour=h, start=5, end=5) for h in (4, 5, 6)]}


def snapshot_round_trip(pkg, tmp):
    L = _m(pkg, "index.local_store")
    SN = _m(pkg, "index.snapshot")
    src = L.LocalStore(tmp / "src.db")
    for i in range(12):
        src.add_document(url=f"https://e.example/{i}", title=f"Title {i}", text=f"tensor memory document number {i} with some body text " * 3, raw_html_hash=f"r{i}", text_hash=f"t{i}",
                         language="en" if i % 2 else None)
    st = SN.export_snapshot(src, tmp / "snap.bin")
    meta = SN.read_snapshot_metadata(tmp / "snap.bin")
    dst = L.LocalStore(tmp / "dst.db")
    dst.add_document(url="https://e.example/3", title="already here", text="a different text entirely", raw_html_hash="x", text_hash="y")
    imp = SN.import_snapshot(dst, tmp / "snap.bin")
    again = SN.import_snapshot(dst, tmp / "snap.bin")
    out = {"export": (st.total_documents, st.exported, st.skipped, st.file_size_bytes > 0), "meta": {k: meta.get(k) for k in ("version", "document_count", "total_documents") if k in meta},
           "import": (imp.total_documents, imp.exported, imp.skipped), "again": (again.exported, again.skipped), "count": dst.get_stats()["document_count"],
           "hit": [r.url for r in dst.search("tensor memory", limit=20)].count("https://e.example/5")}
    src.close()
    dst.close()
    return out


SCENARIOS.update({f.__name__: f for f in (recrawl_selection, subnet_limits, attestations, llm_scheduling, snapshot_round_trip)})


# ----------------------------------------------------------------------------- fifth batch: audits, takedowns, credit proofs
def audit_rounds(pkg, tmp):
    A = _m(pkg, "trust.audit")
    H = _m(pkg, "hashing")
    raw, text = b"<html>body</html>", "body text of the page"
    th, rh = H.content_hash(text), H.content_hash(raw.decode())
    ok = A.perform_audit_check("https://e.example/a", th, rh, actual_raw_body=raw, actual_text=text, auditor_peer_id="aud1", audit_id="x1", target_peer_id="t")
    bad_text = A.perform_audit_check("https://e.example/a", th, rh, actual_raw_body=raw, actual_text=text + "!", auditor_peer_id="aud2", audit_id="x1", target_peer_id="t")
    missing = A.perform_audit_check("https://e.example/a", th, rh, auditor_peer_id="aud3", audit_id="x1", target_peer_id="t")
    sch = A.AuditScheduler()
    t0 = 1_000_000.0
    first = sch.should_schedule(now=t0)
    req = sch.create_audit("target", "https://e.example/a", th, rh, ["a1", "a2", "a3", "a4", "target"], now=t0)
    none = sch.create_audit("target", "https://e.example/b", th, rh, ["target"], now=t0)
    summary = None
    verdicts = []
    if req is not None:
        for i, pid in enumerate(req.auditor_peer_ids):
            res = A.perform_audit_check(req.url, th, rh, actual_raw_body=raw, actual_text=text if i != 1 else "tampered", auditor_peer_id=pid, audit_id=req.audit_id,
                                        target_peer_id=req.target_peer_id)
            verdicts.append(res.verdict.value)
            out = sch.submit_result(res)
            summary = out or summary
    return {"checks": [(r.verdict.value, r.actual_text_hash == th, bool(r.detail)) for r in (ok, bad_text, missing)], "first": first,
            "request": None if req is None else (req.target_peer_id, req.url, len(req.auditor_peer_ids), "target" in req.auditor_peer_ids, len(set(req.auditor_peer_ids))),
            "no_auditors": none is None, "verdicts": sorted(verdicts),
            "summary": None if summary is None else (summary.final_verdict.value, summary.pass_count, summary.fail_count, summary.error_count, len(summary.suspicious_auditors)),
            "canonical_is_bytes": isinstance(A.audit_result_canonical(ok), bytes), "verdict_names": sorted(v.value for v in A.AuditVerdict)}


def takedowns(pkg, tmp):
    D = _m(pkg, "trust.dmca")
    K = _m(pkg, "p2p.keys")
    kp = K.KeyPair.generate()
    tm = D.TakedownManager(str(tmp / "dmca.db"))
    t0 = 1_000_000.0
    n = tm.create_notice("https://e.example/stolen", "copyright: my article", kp, contact_info="me@example.org", now=t0)
    wire = D.serialize_notice(n)
    back = D.deserialize_notice(wire)
    out = {"notice": (n.url, n.reason, n.contact_info, n.created_at, round(n.deadline - n.created_at, 3), len(n.signature), n.requester_id == kp.peer_id),
           "wire_keys": sorted(wire), "round_trip": back == n, "verifies": (tm.verify_notice(back, kp), tm.verify_notice(back, K.KeyPair.generate())),
           "down": (tm.is_taken_down("https://e.example/stolen"), tm.is_taken_down("https://e.example/other")),
           "lookup": (tm.get_notice_for_url("https://e.example/stolen") is not None, tm.get_notice_for_url("https://e.example/other")),
           "status0": tm.check_compliance(n.notice_id, "peerA", now=t0 + 10).value}
    ack = tm.acknowledge(n.notice_id, "peerA", now=t0 + 20)
    out["ack"] = (ack.status.value, ack.peer_id) if ack else None
    out["status1"] = tm.check_compliance(n.notice_id, "peerA", now=t0 + 30).value
    done = tm.mark_complied(n.notice_id, "peerA", now=t0 + 40)
    out["complied"] = (done.status.value, done.complied_at) if done else None
    # (after mark_complied the reference still reports the peer's FIRST acknowledgement -- "acknowledged", and lists the peer as
    #  non-compliant past the deadline; this repo reports "complied": deliberately not compared)
    out["late"] = tm.check_compliance(n.notice_id, "peerB", now=n.deadline + 10).value
    out["non_compliant"] = [x.url for x in tm.list_non_compliant("peerB", now=n.deadline + 10)]
    tm.record_propagation(n.notice_id, "peerC")
    tm.record_propagation(n.notice_id, "peerC")
    rec = tm.get_record(n.notice_id)
    out["record"] = (len(rec.acknowledgments) >= 1, sorted(set(rec.propagated_to))) if rec else None
    out["unknown"] = (tm.acknowledge("nope", "peerA"), tm.get_record("nope"), tm.check_compliance("nope", "peerA").value)
    out["active"] = [x.url for x in tm.list_active()]
    out["dht_key"] = D.takedown_dht_key("https://e.example/stolen")
    tm.close()
    return out


def credit_proofs(pkg, tmp):
    L = _m(pkg, "credits.ledger")
    T = _m(pkg, "credits.types")
    V = _m(pkg, "credits.verification")
    K = _m(pkg, "p2p.keys")
    kp = K.KeyPair.generate()
    led = L.CreditLedger(tmp / "c.db")
    for i in range(12):
        led.record_action(T.ActionType.CRAWL, 1 + i % 3, note=f"n{i}", key_pair=kp)
    led.record_action(T.ActionType.NETWORK_UPTIME, 2.0)            # unsigned entry
    proof = V.CreditProofBuilder(led, kp).build_proof(sample_size=5, request_id="req-1")
    signed = led.signed_entries()
    out = {"keys": sorted(proof), "request_id": proof.get("request_id"), "peer": proof.get("peer_id") == kp.peer_id, "entry_count": proof.get("entry_count"),
           "signed": len(signed), "sample": len(proof.get("sample_entries", proof.get("entries", []))), "total": round(float(proof.get("total_earned", 0.0)), 6)}
    verify = getattr(V, "verify_credit_proof", None)
    if verify is not None:
        res = verify(proof, kp.public_key_bytes()) if "public_key" not in proof else verify(proof)
        out["verified"] = (res.verified, res.invalid_signatures, res.invalid_proofs, res.merkle_root_valid)
    led.close()
    return out


SCENARIOS.update({f.__name__: f for f in (audit_rounds, takedowns, credit_proofs)})


# ----------------------------------------------------------------------------- sixth batch: operations / privacy / reporting
class _ScenarioKey:
    """Deterministic stand-in for an Ed25519 key pair (the managers only need ``peer_id`` / ``sign`` / ``verify``)."""

    def __init__(self, name: str):
        self.peer_id, self._secret = f"peer-{name}", name.encode()

    def sign(self, payload: bytes) -> bytes:
        import hashlib

        return hashlib.sha256(self._secret + payload).digest()

    def verify(self, payload: bytes, signature: bytes) -> bool:
        return signature == self.sign(payload)


def dashboard_formatting(pkg, tmp):
    U = _m(pkg, "dashboard.utils")
    D = _m(pkg, "diagnostics")

    class Tier:
        def __init__(self, name):
            self.name = name

    return {"uptime": [U.format_uptime(s) for s in (-5, 0, 59, 60, 3599, 3600, 86399, 86400, 90061, 10 ** 7)],
            "tier": [U.tier_label(Tier(n)) for n in ("TIER_1", "TIER_2", "TIER_3", "TIER_9", "")] + [U.tier_label(object())],
            "bytes": [U.format_bytes(n) for n in (0, 1, 1023, 1024, 1536, 1024 ** 2, 5 * 1024 ** 3, 1024 ** 5, 3 * 1024 ** 6, -2048, 0.5)],
            "doc": [U.format_doc_line("https://e.com/a", t) for t in ("", "short", "x" * 40, "y" * 41, "z" * 200)],
            "pct": [D._percentile(v, p) for v in ([], [5.0], [3.0, 1.0, 2.0], list(map(float, range(100)))) for p in (0, 50, 95, 99, 100)]}


def partition_detector(pkg, tmp):
    D = _m(pkg, "diagnostics")
    clock = iter(float(1000 + 10 * i) for i in range(1000))
    with mock.patch(f"{pkg}.diagnostics.time.time", side_effect=lambda: next(clock)):
        det = D.PartitionDetector(warning_threshold=0.5, critical_threshold=0.8, min_peers_for_alert=3)
        seen = []
        for count in (10, 10, 10, 9, 4, 4, 1, 1, 12, 12, 12, 2, 0, 0, 20):
            a = det.record(count)
            seen.append(None if a is None else (a.severity, a.previous_peers, a.current_peers, round(a.drop_ratio, 3)) if hasattr(a, "drop_ratio")
                        else (a.severity, a.previous_peers, a.current_peers))
        quiet = D.PartitionDetector(min_peers_for_alert=50)
        calm = [quiet.record(c) is None for c in (10, 10, 10, 1, 0)]
    return {"alerts": seen, "kept": len(det.alerts), "calm": calm}


def gdpr_deletions(pkg, tmp):
    G = _m(pkg, "trust.gdpr")
    alice, bob = _ScenarioKey("alice"), _ScenarioKey("bob")
    db = str(tmp / f"gdpr-{pkg}.db")
    mgr = G.DeletionManager(db)
    r1 = mgr.create_request("https://e.com/me", G.DeletionBasis.ERASURE if hasattr(G.DeletionBasis, "ERASURE") else list(G.DeletionBasis)[0],
                            "my personal page", alice, personal_data_fields=["name", "email"], now=1000.0)
    r2 = mgr.create_request("https://e.com/other", list(G.DeletionBasis)[-1], "old address", alice, now=1001.0)
    out = {"id_len": len(r1.request_id), "ids_differ": r1.request_id != r2.request_id, "requester": r1.requester_id, "fields": list(r1.personal_data_fields),
           "verify": [mgr.verify_request(r1, alice), mgr.verify_request(r1, bob)],
           "blocked": [mgr.is_blocked(u) for u in ("https://e.com/me", "https://e.com/other", "https://e.com/none")],
           "for_url": mgr.get_request_for_url("https://e.com/me").request_id == r1.request_id, "none_for_url": mgr.get_request_for_url("https://x/") is None}
    wire = G.serialize_request(r1)
    back = G.deserialize_request(wire)
    out["wire_keys"] = sorted(wire)
    out["round_trip"] = (back.request_id == r1.request_id, back.url, back.basis.value, back.reason, back.created_at, back.signature == r1.signature)
    peer = G.DeletionManager()
    out["receive"] = [peer.receive_request(back, alice), peer.receive_request(back, alice), peer.receive_request(G.deserialize_request(dict(wire, reason="tampered")), alice)]
    out["peer_blocked"] = peer.is_blocked("https://e.com/me")
    c1 = mgr.confirm_deletion(r1.request_id, "peer-x", now=1010.0)
    c2 = mgr.confirm_deletion("nope", "peer-x", now=1011.0)
    mgr.record_propagation(r1.request_id, "peer-y")
    mgr.record_propagation(r1.request_id, "peer-y")
    rec = mgr.get_record(r1.request_id)
    out["confirm"] = (c1.peer_id, c1.deleted_at, c1.status.value, c2 is None, len(rec.confirmations), sorted(rec.propagated_to))
    out["pending"] = [[r.url for r in mgr.list_pending(p)] for p in ("peer-x", "peer-z")]
    out["all"] = sorted(r.url for r in mgr.list_all())
    out["dht_key"] = (G.deletion_dht_key("https://e.com/me"), G.deletion_dht_key("https://e.com/me") == G.deletion_dht_key("https://e.com/ME"))
    out["unblock"] = [mgr.unblock("https://e.com/other", admin_key=bob), mgr.is_blocked("https://e.com/other"), mgr.unblock("https://e.com/never", admin_key=bob)]
    out["size"] = mgr.blocklist_size
    mgr.close()
    again = G.DeletionManager(db)            # durable: requests, confirmations, propagation and the block list come back
    rec2 = again.get_record(r1.request_id)
    out["reloaded"] = (sorted(r.url for r in again.list_all()), again.is_blocked("https://e.com/me"), again.is_blocked("https://e.com/other"),
                       len(rec2.confirmations), sorted(rec2.propagated_to), [c.status.value for c in rec2.confirmations])
    again.close()
    return out


def llm_reputation(pkg, tmp):
    R = _m(pkg, "trust.reputation")
    t = R.LLMReputationTracker(tmp / f"rep-{pkg}.db")
    for peer, q in (("good", 0.9), ("good", 0.95), ("good", 0.85), ("good", 1.0), ("good", 0.9), ("mid", 0.5), ("mid", 0.6), ("mid", 0.4), ("mid", 0.55),
                    ("mid", 0.5), ("bad", 0.1), ("bad", 0.0), ("bad", 0.2), ("bad", 0.1), ("bad", 0.05), ("new", 0.99), ("weird", 7.0), ("weird", -3.0)):
        t.record_quality(peer, q, url=f"https://e.com/{peer}", content_hash="h")

    def view(r):
        return None if r is None else (r.peer_id, round(r.avg_quality, 4) if hasattr(r, "avg_quality") else None, r.total_summaries if hasattr(r, "total_summaries") else None,
                                       r.grade.value)

    out = {"reps": [view(t.get_reputation(p)) for p in ("good", "mid", "bad", "new", "weird", "absent")],
           "scores": [round(t.get_quality_score(p), 4) for p in ("good", "mid", "bad", "new", "absent")],
           "top": [r.peer_id for r in t.top_peers(3)], "listed": [r.peer_id for r in t.list_peers()],
           "grades": [R._grade_from_score(s, n).value for s, n in ((0.95, 10), (0.95, 1), (0.7, 10), (0.5, 10), (0.3, 10), (0.1, 10), (0.0, 0))]}
    t.close()
    return out


def governor_ladder(pkg, tmp):
    G = _m(pkg, "resources.governor")
    P = _m(pkg, "resources.profiles")
    out = []
    for profile in ("minimal", "balanced", "dedicated"):
        gov = G.ResourceGovernor(P.get_profile(profile))
        for cpu, mem, rss in ((5, 10, 10), (45, 50, 100), (55, 50, 100), (65, 50, 100), (75, 60, 100), (82, 60, 100), (50, 86, 100), (92, 60, 100), (50, 50, 480),
                              (50, 50, 520), (97, 50, 100), (50, 50, 700), (10, 10, 10)):
            with mock.patch.object(G.ResourceGovernor, "_sample_cpu", staticmethod(lambda cpu=cpu: float(cpu))), \
                    mock.patch.object(G.ResourceGovernor, "_sample_memory", staticmethod(lambda mem=mem: float(mem))), \
                    mock.patch.object(G.ResourceGovernor, "_sample_process_memory_mb", lambda self, rss=rss: float(rss)):
                st = gov.check_and_adjust()
            out.append((profile, cpu, mem, rss, int(st.degrade_level), round(st.throttle_factor, 3), gov.should_throttle_crawl, gov.should_pause_crawl,
                        gov.should_disable_llm, gov.should_disable_remote_search, gov.is_read_only, gov.effective_max_concurrent))
    return out


def score_explanations(pkg, tmp):
    E = _m(pkg, "search.explain")
    RR = _m(pkg, "index.ranking").RankedResult
    hits = [RR(doc_id=i, url=f"https://e.com/{i}", title=f"T{i}", snippet="s", bm25_score=b, freshness_score=f, trust_score=t, authority_score=a,
               combined_score=c, crawled_at=1.0, peer_id="") for i, (b, f, t, a, c) in enumerate(((0.9, 0.5, 0.5, 0.2, 0.61), (0.0, 0.0, 0.0, 0.0, 0.0),
                                                                                                 (1.0, 1.0, 1.0, 1.0, 1.0), (0.3, 0.9, 0.1, 0.0, 0.33)))]
    def contract(d):          # this package adds `breakdown` / `dominant_factor` for the MCP explain tool: a superset, compared without them
        return {k: v for k, v in d.items() if k not in ("breakdown", "dominant_factor")}

    single = [contract(E.explain_result(h).to_dict()) for h in hits]
    whole = E.explain_query("How  to AND x", "how to x", hits, 12.3456, pipeline=["sanitize", "fts5", "rank"]).to_dict()
    whole["results"] = [contract(r) for r in whole["results"]]
    bare = E.explain_query("q", "q", [], 0.0).to_dict()
    return {"single": single, "whole": whole, "bare": bare}


SCENARIOS.update({f.__name__: f for f in (dashboard_formatting, partition_detector, gdpr_deletions, llm_reputation, governor_ladder, score_explanations)})


# ----------------------------------------------------------------------------- seventh batch: wire envelopes, topology helpers, seeds
def signed_envelopes(pkg, tmp):
    """Both packages sign with a real Ed25519 key; canonical bytes, the wire dict and every rejection reason must agree, and an
    envelope signed by one package must verify in the other (the caller compares ``cross``)."""
    A = _m(pkg, "p2p.message_auth")
    from cryptography.hazmat.primitives.asymmetric.ed25519 import Ed25519PrivateKey
    from cryptography.hazmat.primitives.serialization import Encoding, PublicFormat

    class Key:
        def __init__(self, seed: bytes, peer_id: str):
            self._k, self.peer_id = Ed25519PrivateKey.from_private_bytes(seed), peer_id

        def sign(self, data: bytes) -> bytes:
            return self._k.sign(data)

        def public_key_bytes(self) -> bytes:
            return self._k.public_key().public_bytes(Encoding.Raw, PublicFormat.Raw)

    alice, mallory = Key(b"a" * 32, "12D3KooWAlice"), Key(b"m" * 32, "12D3KooWMallory")
    out = {"canonical": [A._canonical_bytes(p, n, t, b).hex() for p, n, t, b in (("peer", 0, 0.0, b""), ("12D3KooWAlice", 7, 1700000000.123456789, b"\x01\x02"),
                                                                                   ("p|q", 2 ** 40, 1.5, b"|"))]}
    counter = A.NonceCounter()
    env = [A.sign_envelope(bytes([i]) * 3, alice, counter, now=1000.0 + i) for i in range(4)]
    out["nonces"] = [e.nonce for e in env] + [counter.current]
    wire = A.envelope_to_dict(env[0])
    out["wire"] = {k: (v.hex() if isinstance(v, bytes) else v) for k, v in sorted(wire.items())}
    back = A.envelope_from_dict(wire)
    out["round_trip"] = (back.payload, back.peer_id, back.nonce, back.timestamp, back.signature == env[0].signature)
    reg, seen = A.PeerKeyRegistry(), A.NonceTracker()
    reg.register(alice.peer_id, alice.public_key_bytes())

    def attempt(e, **kw):
        try:
            return ("ok", A.verify_envelope(e, reg, seen, **kw).hex())
        except A.VerificationError as exc:
            return ("rejected", str(exc))

    import dataclasses

    forged = dataclasses.replace(env[2], payload=b"evil")
    stranger = A.sign_envelope(b"x", mallory, A.NonceCounter(), now=1000.0)
    # (an OLDER nonce arriving late is a deliberate difference: the reference demands strictly increasing nonces, this package
    #  accepts anything inside a sliding window that it has not seen -- tests/test_p2p_* cover that; not compared here)
    out["verify"] = [attempt(env[0], now=1001.0), attempt(env[1], now=1001.0), attempt(env[1], now=1001.0), attempt(forged, now=1002.0),
                     attempt(env[3], now=1003.0 + 10_000), attempt(stranger, now=1000.0), attempt(env[3], now=1003.0, is_isolated_fn=lambda p: True),
                     attempt(env[3], now=1003.0)]
    out["highest"] = (seen.highest(alice.peer_id), seen.highest("nobody"), alice.peer_id in reg, len(reg))
    reg.remove(alice.peer_id)
    out["after_remove"] = (alice.peer_id in reg, reg.get(alice.peer_id))
    return out


def topology_helpers(pkg, tmp):
    N = _m(pkg, "p2p.network_ext")
    G = N.GeoLocation
    seoul, tokyo, paris, origin = G("KR", "", "Seoul", 37.5665, 126.978), G("JP", "", "Tokyo", 35.6762, 139.6503), G("FR", "", "Paris", 48.8566, 2.3522), G()
    out = {"km": [round(N.estimate_geo_distance(a, b), 3) for a, b in ((seoul, tokyo), (tokyo, seoul), (seoul, paris), (seoul, seoul), (origin, paris))],
           "near": [(p, round(d, 1)) for p, d in N.sort_peers_by_proximity([("paris", paris), ("tokyo", tokyo), ("here", seoul), ("null", origin)], seoul)],
           "relay": [N.select_relay(list(r)) for r in ([], [("a", 50.0)], [("slow", 200.0), ("fast", 20.0), ("mid", 90.0)], [("x", 5.0), ("y", 5.0)])]}
    det = N.PartitionDetector(threshold=0.5)
    steps = []
    with mock.patch(f"{pkg}.p2p.network_ext.time.time", return_value=123.0):
        # (one uninterrupted partition: this package resets the attempt counter once the partition heals, the reference never does)
        for reachable, total in ((10, 10), (5, 10), (0, 0), (4, 10), (1, 10), (1, 10), (1, 10), (1, 10), (1, 10), (9, 10)):
            st = det.check(reachable, total)
            steps.append((st.is_partitioned, st.reachable_peers, st.expected_peers, det.get_recovery_actions()))
    out["partition"] = steps
    return out


def seeds_and_identity(pkg, tmp):
    S = _m(pkg, "crawler.seeds")
    I = _m(pkg, "credits.github_identity")
    d = tmp / f"seeds-{pkg}"
    d.mkdir()
    cats = list(S.CATEGORIES)
    (d / f"{cats[0]}.txt").write_text("# comment\nhttps://a.example/\n\n  https://b.example/x  \nftp://nope.example\nnot a url\nhttp://c.example\n")
    (d / f"{cats[1]}.txt").write_text("https://z.example/\n#https://commented.example\n")
    (d / "zz-extra.txt").write_text("https://extra.example/\n")
    (d / "ignored.md").write_text("https://md.example/\n")
    return {"categories": sorted(cats), "one": S.load_seeds(cats[0], d), "two": S.load_seeds(cats[1], d), "missing_file": S.load_seeds(cats[2], d) if len(cats) > 2 else [],
            "unknown": S.load_seeds("no-such-category", d), "all": sorted(S.load_seeds(None, d)), "no_dir": S.load_seeds(None, d / "absent"),
            "bundled_nonempty": len(S.load_seeds()) > 0,
            "emails": [I.is_valid_email(e) for e in ("a@b.co", "first.last+tag@sub.example.org", "no-at.example", "a@b", "@b.co", "a b@c.de", "", "a@b.c")],
            "msg": [I.format_startup_message("me@example.org"), I.format_startup_message(None), I.format_startup_message("")]}


SCENARIOS.update({f.__name__: f for f in (signed_envelopes, topology_helpers, seeds_and_identity)})


# ----------------------------------------------------------------------------- eighth batch: LLM glue, WET import, preflight
class _ScriptedLLM:
    """An LLM backend that answers from a script and records the prompts it was given."""

    def __init__(self, pkg, answers):
        E = _m(pkg, "summarizer.engine")
        self._E, self.answers, self.prompts = E, list(answers), []

    async def generate(self, prompt: str, *, max_tokens: int = 512) -> str:
        self.prompts.append((prompt, max_tokens))
        a = self.answers.pop(0)
        if isinstance(a, Exception):
            raise a
        return a

    async def is_available(self) -> bool:
        return True

    async def model_info(self):
        return self._E.ModelInfo(name="scripted", runtime=self._E.LLMRuntime.OLLAMA, parameter_count="0", quantization="none", available=True) \
            if "parameter_count" in self._E.ModelInfo.__dataclass_fields__ else self._E.ModelInfo(**{f: ("scripted" if f == "name" else self._E.LLMRuntime.OLLAMA if f == "runtime" else None)
                                                                                                  for f in self._E.ModelInfo.__dataclass_fields__})


def llm_rerank_and_summaries(pkg, tmp):
    R = _m(pkg, "search.reranker")
    E = _m(pkg, "summarizer.engine")
    RR = _m(pkg, "index.ranking").RankedResult
    hits = [RR(doc_id=i, url=f"https://e.com/{i}", title=f"Title {i}", snippet=("snippet %d " % i) * 30, bm25_score=0.5, freshness_score=0.5, trust_score=0.5,
               authority_score=0.5, combined_score=1.0 - i / 10, crawled_at=1.0, peer_id="") for i in range(6)]
    out = {"block": R._build_results_block(hits[:3], max_snippet=40),
           "parse": [R._parse_ranking_response(t, 4) for t in ("3,1,2,4", "[2, 4, 1, 3]", "Ranking: 4 > 2 > 1 > 3", "1, 1, 2", "2,9,1", "", "no digits here",
                                                                "3\n1\n2\n4", "1. 4\n2. 3", "4,3")]}

    async def go():
        orders = []
        for answer, kw in (("3,1,2", dict(max_candidates=3)), ("2,1", dict(max_candidates=3)), ("garbage", dict(max_candidates=3)), (RuntimeError("down"), dict(max_candidates=3)),
                           ("4,3,2,1", dict(max_candidates=4, top_n=2)), ("1", dict(max_candidates=1))):
            llm = _ScriptedLLM(pkg, [answer])
            res = await R.rerank_with_llm("what is tcgen05", hits, llm, **kw)
            orders.append(([h.doc_id for h in res], len(llm.prompts), llm.prompts[0][1] if llm.prompts else None))
        empty = await R.rerank_with_llm("q", [], _ScriptedLLM(pkg, []))
        llm = _ScriptedLLM(pkg, ["  A short summary.  ", "", RuntimeError("boom")])
        eng = E.SummarizationEngine(llm)
        s1 = await eng.summarize("https://e.com/doc", "Doc title", "body text " * 2000, max_tokens=64, max_input_chars=500)
        s2 = await eng.summarize("https://e.com/doc2", "", "tiny", max_tokens=32)
        try:
            s3 = await eng.summarize("https://e.com/doc3", "T", "text")
            third = ("ok", s3.summary)
        except Exception as exc:  # noqa: BLE001
            third = ("raised", type(exc).__name__)
        return orders, [h.doc_id for h in empty], (s1.url, s1.summary, s1.model, str(s1.runtime), s1.content_hash, s1.token_count), \
            (s2.summary, s2.content_hash, s2.token_count), third, [(len(p), mt, p[:60], "body text body text" in p) for p, mt in llm.prompts], \
            [E._estimate_tokens(t) for t in ("", "one", "four words right here", "x" * 400)]

    out["orders"], out["empty"], out["summary1"], out["summary2"], out["summary3"], out["prompts"], out["tokens"] = asyncio.run(go())
    return out


WET = ("WARC/1.0\r\nWARC-Type: warcinfo\r\nContent-Length: 10\r\n\r\nsoftware x\r\n\r\n"
       "WARC/1.0\r\nWARC-Type: conversion\r\nWARC-Target-URI: https://a.example/one\r\nWARC-Date: 2024-01-02T03:04:05Z\r\nContent-Length: 400\r\n\r\n" + "alpha beta gamma " * 30 + "\r\n\r\n"
       "WARC/1.0\r\nWARC-Type: conversion\r\nWARC-Target-URI: https://a.example/short\r\nWARC-Date: 2024-01-02T03:04:06Z\r\nContent-Length: 5\r\n\r\ntiny\r\n\r\n"
       "WARC/1.0\r\nWARC-Type: conversion\r\nWARC-Date: 2024-01-02T03:04:07Z\r\nContent-Length: 20\r\n\r\nno target uri here\r\n\r\n"
       "WARC/1.0\r\nWARC-Type: conversion\r\nWARC-Target-URI: https://b.example/two\r\nContent-Length: 300\r\n\r\n" + "delta epsilon zeta " * 20 + "\r\n\r\n"
       "WARC/1.0\r\nWARC-Type: conversion\r\nWARC-Target-URI: https://a.example/one\r\nWARC-Date: 2024-02-02T00:00:00Z\r\nContent-Length: 400\r\n\r\n" + "alpha beta gamma " * 30 + "\r\n\r\n")


def wet_import(pkg, tmp):
    C = _m(pkg, "index.commoncrawl")
    LS = _m(pkg, "index.local_store").LocalStore
    DD = _m(pkg, "crawler.dedup").DeduplicatorDB
    recs = C.parse_wet_content(WET)
    out = {"records": [(r.url, len(r.text), r.date, r.content_length) for r in recs], "none": C.parse_wet_content(""), "junk": C.parse_wet_content("not a warc file at all")}
    wet_path = tmp / f"sample-{pkg}.wet"
    wet_path.write_text(WET)
    import gzip

    gz_path = tmp / f"sample-{pkg}.wet.gz"
    gz_path.write_bytes(gzip.compress(WET.encode()))
    urls = tmp / f"urls-{pkg}.txt"
    urls.write_text("https://a.example/one\n# c\nhttps://new.example/\n\nnot-a-url\nhttps://new.example/\nhttps://third.example/p?utm_source=x\n")
    store, dedup = LS(tmp / f"cc-{pkg}.db"), DD(str(tmp / f"cc-dedup-{pkg}.db"))

    async def go():
        imp = C.CommonCrawlImporter(store, dedup)
        a = await imp.import_wet_file(str(wet_path))
        b = await imp.import_wet_file(str(gz_path))                  # everything is a duplicate now
        c = await imp.import_url_list(str(urls), max_urls=10)
        d = await imp.import_url_list(str(urls), max_urls=10)
        try:
            await imp.import_wet_file(str(tmp / "missing.wet"))
            missing = "no error"
        except Exception as exc:  # noqa: BLE001
            missing = type(exc).__name__
        return [(s.total_records, s.imported, s.skipped_duplicate, s.skipped_too_short, s.skipped_error) for s in (a, b, c, d)], missing

    out["stats"], out["missing"] = asyncio.run(go())
    out["stored"] = [store.get_document_by_url(u) is not None for u in ("https://a.example/one", "https://b.example/two", "https://a.example/short",
                                                                          "https://new.example/")]
    out["count"] = store.get_stats()["document_count"]
    store.close()
    dedup.close()
    return out


def preflight_checks(pkg, tmp):
    P = _m(pkg, "resources.preflight")
    out = {}
    for free in (10_000.0, 1_500.0, 600.0, 450.0, 250.0, 150.0, 0.0):
        with mock.patch(f"{pkg}.resources.preflight.get_disk_free_mb", return_value=free):
            out[str(free)] = ([(i.severity.value, i.check) for i in P.check_disk_space(tmp)], P.is_disk_critically_low(tmp))
    out["real_positive"] = P.get_disk_free_mb(tmp) > 0
    return out


SCENARIOS.update({f.__name__: f for f in (llm_rerank_and_summaries, wet_import, preflight_checks)})


# ----------------------------------------------------------------------------- ninth batch: the DHT facade over an in-memory Kademlia
class _MemoryKad:
    """``put_value`` / ``get_value`` of a Kademlia node, on a dict; can be told to fail."""

    def __init__(self):
        self.data, self.fail = {}, False

    async def put_value(self, key, value):
        if self.fail:
            raise ConnectionError("no peers")
        self.data[key] = value

    async def get_value(self, key):
        if self.fail:
            raise ConnectionError("no peers")
        return self.data.get(key)


def dht_facade(pkg, tmp):
    D = _m(pkg, "p2p.dht")
    kad = _MemoryKad()
    me, other = D.InfoMeshDHT(kad, "peer-me"), D.InfoMeshDHT(kad, "peer-other")
    ptr = lambda peer, doc, score: {"peer_id": peer, "doc_id": doc, "url": f"https://e.com/{doc}", "score": score, "title": f"T{doc}"}  # noqa: E731

    async def go():
        out = {}
        clock = [1000.0]
        with mock.patch(f"{pkg}.p2p.dht.time.time", side_effect=lambda: clock[0]):
            out["publish"] = [await me.publish_keyword("tensor", [ptr("peer-me", 1, 0.9), ptr("peer-me", 2, 0.5)]),
                              await other.publish_keyword("tensor", [ptr("peer-other", 1, 0.7), ptr("peer-me", 1, 0.1), {"peer_id": 5, "doc_id": "x"}]),
                              await me.publish_keyword("", []), await me.publish_keyword("empty", [])]
            got = await me.query_keyword("tensor")
            out["query"] = sorted((p["peer_id"], p["doc_id"], p["score"]) for p in got)
            out["query_none"] = await me.query_keyword("absent")
            # per-keyword publish rate limit: hammer one keyword inside a minute, then let the window slide
            burst = []
            for i in range(14):
                clock[0] += 1.0
                burst.append(await me.publish_keyword("hot", [ptr("peer-me", 100 + i, 0.1)]))
            clock[0] += 3700.0
            burst.append(await me.publish_keyword("hot", [ptr("peer-me", 999, 0.1)]))
            out["burst"] = burst
            out["hot_count"] = len(await me.query_keyword("hot"))
            # crawl locks: exclusive within the TTL, re-entrant for the holder, released only by the holder, expiring
            clock[0] = 5000.0
            locks = [await me.acquire_crawl_lock("https://e.com/page"), await other.acquire_crawl_lock("https://e.com/page"), await me.acquire_crawl_lock("https://e.com/page"),
                     await other.release_crawl_lock("https://e.com/page"), await other.acquire_crawl_lock("https://e.com/page")]
            clock[0] += 301.0
            locks.append(await other.acquire_crawl_lock("https://e.com/page"))            # expired -> taken over
            locks += [await me.release_crawl_lock("https://e.com/page"), await other.release_crawl_lock("https://e.com/page"),
                      await me.acquire_crawl_lock("https://e.com/page"), await me.release_crawl_lock("https://e.com/never-locked")]
            out["locks"] = locks
            out["attest"] = [await me.publish_attestation("https://e.com/a", "rawhash", "texthash", b"sig")]
            att = await other.get_attestation("https://e.com/a")
            out["attestation"] = {k: (v.hex() if isinstance(v, bytes) else v) for k, v in sorted((att or {}).items()) if k not in ("timestamp", "published_at")}
            out["no_attestation"] = await other.get_attestation("https://e.com/none")
            out["raw"] = [await me.put("/custom/key", b"value"), await other.get("/custom/key"), await other.get("/custom/absent")]
            kad.fail = True
            out["offline"] = [await me.publish_keyword("tensor", [ptr("peer-me", 3, 0.3)]), await me.query_keyword("tensor"), await me.acquire_crawl_lock("https://e.com/x"),
                              await me.release_crawl_lock("https://e.com/page"), await me.publish_attestation("https://e.com/b", "r", "t"),
                              await me.get_attestation("https://e.com/a"), await me.put("/k", b"v"), await me.get("/custom/key")]
        st = me.stats
        out["stats"] = {f: getattr(st, f) for f in sorted(vars(st)) if isinstance(getattr(st, f), (int, float)) and "time" not in f and "ms" not in f}
        out["merge"] = D._merge_pointers([ptr("a", 1, 0.1), ptr("b", 1, 0.2), {"peer_id": None, "doc_id": 1}], [ptr("a", 1, 0.9), ptr("c", 2, 0.3)], limit=3)
        return out

    return asyncio.run(go())


SCENARIOS.update({f.__name__: f for f in (dht_facade,)})


# ----------------------------------------------------------------------------- tenth batch: configuration, throttling, compression
CONFIG_TOML = """
[node]
role = "search"
log_level = "LOUD"
listen_port = 70000
[crawl]
max_concurrent = 0
politeness_delay = 0.01
urls_per_hour = 500
user_agent = "CustomBot/1.0"
[network]
upload_limit_mbps = 2000.0
bootstrap_nodes = ["/ip4/10.0.0.1/tcp/4001/p2p/12D3KooWA", "/ip4/10.0.0.2/tcp/4001/p2p/12D3KooWB"]
replication_factor = 5
[index]
fts_tokenizer = "klingon"
[storage]
compression_level = 99
[resources]
profile = "dedicated"
cpu_nice = -5
[dashboard]
theme = "dracula"
refresh_interval = 60.0
[unknown_section]
x = 1
"""


def config_loading(pkg, tmp):
    C = _m(pkg, "config")
    import os

    path = tmp / f"config-{pkg}.toml"
    path.write_text(CONFIG_TOML)
    env = {"INFOMESH_CRAWL_MAX_DEPTH": "7", "INFOMESH_NODE_LISTEN_PORT": "4100", "INFOMESH_LLM_ENABLED": "yes", "INFOMESH_NETWORK_BOOTSTRAP_NODES": "/ip4/1.1.1.1/tcp/1 , ,/ip4/2.2.2.2/tcp/2",
           "INFOMESH_STORAGE_COMPRESSION_ENABLED": "0", "INFOMESH_NODE_DATA_DIR": str(tmp / "data")}
    with mock.patch.dict(os.environ, env, clear=False):
        cfg = C.load_config(path)
    plain = C.load_config(tmp / "does-not-exist.toml")

    def view(c):
        return {"role": c.node.role, "log_level": c.node.log_level, "port": c.node.listen_port, "data_dir_name": c.node.data_dir.name,
                "max_concurrent": c.crawl.max_concurrent, "delay": c.crawl.politeness_delay, "per_hour": c.crawl.urls_per_hour, "depth": c.crawl.max_depth, "ua": c.crawl.user_agent,
                "up": c.network.upload_limit_mbps, "boot": list(c.network.bootstrap_nodes)[:3], "repl": c.network.replication_factor, "tok": c.index.fts_tokenizer,
                "zlevel": c.storage.compression_level, "zon": c.storage.compression_enabled, "profile": c.resources.profile, "nice": c.resources.cpu_nice,
                "theme": c.dashboard.theme, "refresh": c.dashboard.refresh_interval, "llm": c.llm.enabled, "db_under_data": c.index.db_path.parent == c.node.data_dir}

    out = {"loaded": view(cfg), "defaults": {k: v for k, v in view(plain).items() if k not in ("boot", "data_dir_name")}}
    # out-of-range numbers are clamped, unknown enum strings fall back to the default: observed through the loader
    probes = {}
    for section, key, raw in (("node", "listen_port", "0"), ("node", "listen_port", "8080"), ("crawl", "politeness_delay", "100.0"), ("resources", "cpu_nice", "25"),
                              ("node", "role", "FULL"), ("node", "role", "boss"), ("dashboard", "theme", "nord"), ("dashboard", "theme", "neon"),
                              ("storage", "compression_level", "3"), ("llm", "enabled", "TRUE"), ("llm", "enabled", "0"), ("crawl", "max_depth", "12")):
        with mock.patch.dict(os.environ, {f"INFOMESH_{section.upper()}_{key.upper()}": raw}, clear=False):
            got = getattr(getattr(C.load_config(tmp / "does-not-exist.toml"), section), key)
        probes[f"{section}.{key}={raw}"] = got
    out["probes"] = probes
    saved = tmp / f"saved-{pkg}" / "config.toml"
    C.save_config(cfg, saved)
    again = C.load_config(saved)
    out["round_trip"] = view(again) == view(cfg)
    out["saved_has_sections"] = {"crawl", "node", "network"} <= {line.strip("[]\n") for line in saved.read_text().splitlines() if line.startswith("[")}
    return out


def bandwidth_throttle(pkg, tmp):
    T = _m(pkg, "p2p.throttle")
    clock = [100.0]
    slept = []

    async def fake_sleep(seconds):
        slept.append(round(seconds, 6))
        clock[0] += seconds

    async def go():
        with mock.patch(f"{pkg}.p2p.throttle.time.monotonic", side_effect=lambda: clock[0]), mock.patch(f"{pkg}.p2p.throttle.asyncio.sleep", fake_sleep):
            bucket = T.BandwidthBucket(8.0)                  # 1 MB/s, starts with one second of burst
            waits = [round(await bucket.acquire(n), 6) for n in (0, -5, 400_000, 600_000, 500_000, 2_500_000)]
            clock[0] += 10.0                                  # idle: refills, but never past one second of burst
            waits.append(round(await bucket.acquire(1_000_000), 6))
            waits.append(round(await bucket.acquire(250_000), 6))
            free = T.BandwidthBucket(0.0)
            waits.append(await free.acquire(10 ** 9))
            thr = T.BandwidthThrottle(upload_mbps=8.0, download_mbps=0.8)
            up = [round(await thr.acquire_upload(n), 6) for n in (1_000_000, 500_000)]
            down = [round(await thr.acquire_download(n), 6) for n in (100_000, 50_000, 0)]
            st = thr.stats
            return waits, up, down, (st.upload_bytes, st.download_bytes, st.upload_waits, st.download_waits), bucket.rate_bytes_per_sec

    waits, up, down, stats, rate = asyncio.run(go())
    # (how a long wait is cut into sleeps is an implementation detail -- the reference naps once per second of burst, this
    #  package once per request; the total per call is what callers see)
    return {"waits": waits, "up": up, "down": down, "stats": stats, "rate": rate, "total_sleep": round(sum(slept), 6)}


def compression_round_trips(pkg, tmp):
    Z = _m(pkg, "compression.zstd")
    text = "InfoMesh stores page text compressed. " * 200
    c = Z.Compressor(level=3)
    blob = c.compress(text.encode())
    out = {"level": c.level, "smaller": len(blob) < len(text) // 5, "frame_magic": blob[:4].hex(), "round": c.decompress(blob) == text.encode(),
           "text_round": c.decompress_text(c.compress_text(text)) == text, "empty": c.decompress(c.compress(b"")), "unicode": c.decompress_text(c.compress_text("한국어 텍스트 ✓")),
           "levels": [Z.Compressor(level=lv).level for lv in (1, 9, 19)]}
    for name, call in (("garbage", lambda: c.decompress(b"definitely not zstd")), ("bomb", lambda: c.decompress(blob, max_output_size=100))):
        try:
            call()
            out[name] = "no error"
        except Exception as exc:  # noqa: BLE001
            out[name] = "raised"
    other = Z.Compressor(level=19)
    out["cross_level"] = other.decompress(blob) == text.encode()
    samples = [(f"https://example.org/page/{i} " + "common boilerplate header footer navigation " * 5 + str(i) * 20).encode() for i in range(200)]
    try:
        d = Z.train_dictionary(samples, dict_size=4096)
        with_dict = Z.Compressor(level=3, dict_data=d)
        small = samples[7]
        out["dictionary"] = (len(d) > 0, with_dict.decompress(with_dict.compress(small)) == small)
    except Exception as exc:  # noqa: BLE001
        out["dictionary"] = ("raised", type(exc).__name__)
    return out


SCENARIOS.update({f.__name__: f for f in (config_loading, bandwidth_throttle, compression_round_trips)})


# ----------------------------------------------------------------------------- eleventh batch: peer summarisation service, bootstrap helpers
def peer_summarisation_service(pkg, tmp):
    H = _m(pkg, "summarizer.peer_handler")
    E = _m(pkg, "summarizer.engine")
    clock = [10_000.0]

    async def go():
        with mock.patch(f"{pkg}.summarizer.peer_handler.time.time", side_effect=lambda: clock[0]):
            llm = _ScriptedLLM(pkg, ["Summary one.", RuntimeError("model crashed"), "Summary three.", "Summary four."])
            h = H.PeerSummarizationHandler(E.SummarizationEngine(llm), min_trust_score=0.3)

            def req(i, peer="peer-a", text="body " * 50):
                return H.SummarizeRequest(request_id=f"r{i}", requester_peer_id=peer, url=f"https://e.com/{i}", title=f"T{i}", text=text, max_tokens=64, timestamp=clock[0])

            seq = []
            for r, trust, dt in ((req(1), 0.9, 0), (req(2), 0.9, 1), (req(3), 0.1, 100), (req(4, text="x" * 20_000), 0.9, 100), (req(5), 0.9, 100), (req(6, peer="peer-b"), 0.9, 0),
                                 (req(7), 0.9, 200)):
                clock[0] += dt
                resp = await h.handle_request(r, requester_trust=trust)
                seq.append((resp.request_id, resp.status.value, resp.summary, resp.reject_reason.value if resp.reject_reason else None, resp.model, bool(resp.content_hash)))
            counters = (h.active_count, h.total_served, h.total_rejected)
            wire_req = H.serialize_request(req(9))
            back_req = H.deserialize_request(wire_req)
            ok = H.SummarizeResponse(request_id="r9", status=H.RequestStatus.COMPLETED if hasattr(H.RequestStatus, "COMPLETED") else list(H.RequestStatus)[0], summary="s",
                                     content_hash="h", model="m", elapsed_ms=12.5)
            no = H.SummarizeResponse(request_id="r10", status=list(H.RequestStatus)[-1], reject_reason=list(H.RejectReason)[0], detail="why")
            wires = [H.serialize_response(ok), H.serialize_response(no)]
            backs = [H.deserialize_response(w) for w in wires]
            return {"sequence": seq, "counters": counters, "statuses": [s.value for s in H.RequestStatus], "reasons": [r.value for r in H.RejectReason],
                    "wire_req": {k: v for k, v in sorted(wire_req.items())}, "req_round": (back_req.request_id, back_req.requester_peer_id, back_req.url, back_req.max_tokens, back_req.timestamp),
                    "wire_resp": [{k: v for k, v in sorted(w.items())} for w in wires],
                    "resp_round": [(b.request_id, b.status.value, b.summary, b.reject_reason.value if b.reject_reason else None, b.detail, b.elapsed_ms) for b in backs],
                    "limits": (H.MAX_PENDING_PER_PEER, H.MAX_CONCURRENT_REQUESTS, H.MAX_TEXT_LENGTH, H.PEER_COOLDOWN_SECONDS, H.PROTOCOL_LLM)}

    return asyncio.run(go())


def bootstrap_helpers(pkg, tmp):
    B = _m(pkg, "p2p.bootstrap")
    nodes = B.discover_from_static([{"addr": "/ip4/203.0.113.5/tcp/4001/p2p/12D3KooWA", "region": "eu"}, {"addr": "/dns4/boot.example.org/tcp/4002"}, {"region": "no addr"},
                                    "not a dict", {"addr": "/ip6/2001:db8::1/tcp/abc"}, {"addr": "garbage"}])
    out = {"static": [(n.addr, n.source, n.region, n.host_port) for n in nodes]}
    with mock.patch(f"{pkg}.p2p.bootstrap.time.time", return_value=1_000_000.0):
        peers = [{"peer_id": "fresh-short", "addr": "a", "last_seen": 999_990.0, "uptime": 60}, {"peer_id": "fresh-long", "addr": "b", "last_seen": 999_000.0, "uptime": 90_000},
                 {"peer_id": "stale-long", "addr": "c", "last_seen": 800_000.0, "uptime": 500_000}, {"peer_id": "junk", "addr": "d", "last_seen": "yesterday", "uptime": None},
                 {"peer_id": "mid", "addr": "e", "last_seen": 960_000.0, "uptime": 40_000}]
        out["seeds"] = [[p["peer_id"] for p in B.select_seed_peers(list(peers), max_peers=k)] for k in (2, 10)]
        out["no_seeds"] = B.select_seed_peers([], max_peers=3)
    clock = [0.0]
    with mock.patch(f"{pkg}.p2p.bootstrap.time.time", side_effect=lambda: clock[0]), mock.patch(f"{pkg}.p2p.bootstrap.time.monotonic", side_effect=lambda: clock[0]):
        rl = B.BootstrapRateLimiter(max_per_minute=3, window_seconds=60.0)
        seq = []
        for who, dt in (("a", 0), ("a", 1), ("a", 1), ("a", 1), ("b", 0), ("a", 58), ("a", 1), ("a", 0)):
            clock[0] += dt
            seq.append(rl.allow(who))
        out["rate"] = seq
        out["tracked"] = rl.tracked_clients
        rl.reset("a")
        out["after_reset"] = (rl.allow("a"), rl.tracked_clients)
        clock[0] += 1000
        out["cleanup"] = (rl.cleanup(), rl.tracked_clients)
    return out


SCENARIOS.update({f.__name__: f for f in (peer_summarisation_service, bootstrap_helpers)})


# ----------------------------------------------------------------------------- twelfth batch: robots.txt policy, resource profiles
ROBOTS = {
    "https://open.example": (200, "User-agent: *\nAllow: /\nSitemap: https://open.example/sitemap.xml\nsitemap: https://open.example/news.xml\n"),
    "https://strict.example": (200, "User-agent: *\nDisallow: /private/\nDisallow: /tmp\nCrawl-delay: 7\n\nUser-agent: InfoMeshBot\nDisallow: /bots-only/\nCrawl-delay: 2.5\n"),
    "https://closed.example": (200, "User-agent: *\nDisallow: /\n"),
    "https://missing.example": (404, "not found"),
    "https://broken.example": (500, "oops"),
    "https://weird.example": (200, "this is not\na robots file <html>\nCrawl-delay: soon\n"),
}


class _RobotsClient:
    def __init__(self):
        self.calls = []

    async def get(self, url, **kw):
        import httpx

        self.calls.append(url)
        base = url.rsplit("/robots.txt", 1)[0]
        if base == "https://down.example":
            raise httpx.ConnectError("refused")
        status, text = ROBOTS.get(base, (404, ""))

        class R:
            status_code, = (status,)

        R.text = text
        return R()


def robots_policy(pkg, tmp):
    R = _m(pkg, "crawler.robots")

    async def go():
        chk, client = R.RobotsChecker("InfoMeshBot/1.0", cache_ttl=3600), _RobotsClient()
        urls = ["https://open.example/a", "https://strict.example/private/x", "https://strict.example/public", "https://strict.example/tmp/file", "https://strict.example/bots-only/y",
                "https://closed.example/", "https://closed.example/anything", "https://missing.example/x", "https://broken.example/x", "https://down.example/x",
                "https://weird.example/page", "https://open.example/b"]
        try:
            allowed = [await chk.is_allowed(client, u) for u in urls]
        except Exception as exc:  # noqa: BLE001 -- the reference parses `Crawl-delay: soon` with float(): compared as an outcome
            allowed = ("raised", type(exc).__name__)
        doms = ("open.example", "strict.example", "closed.example", "missing.example", "down.example", "never-seen.example")
        out = {"allowed": allowed, "fetches": len(client.calls), "sitemaps": [chk.get_sitemaps(d) for d in doms], "delays": [chk.get_crawl_delay(d) for d in doms]}
        chk.clear_cache()
        before = len(client.calls)
        await chk.is_allowed(client, "https://open.example/c")
        out["refetched_after_clear"] = len(client.calls) - before
        return out

    return asyncio.run(go())


def resource_profiles(pkg, tmp):
    P = _m(pkg, "resources.profiles")
    ref_fields = ("cpu_cores_limit", "cpu_nice", "memory_limit_mb", "disk_io_priority", "upload_limit_mbps", "download_limit_mbps", "max_concurrent_crawl", "llm_enabled",
                  "llm_off_peak_only")

    def view(p):
        return {f: getattr(p, f) for f in ref_fields} | {"name": str(p.name)}

    out = {"named": [view(P.get_profile(n)) for n in ("minimal", "balanced", "contributor", "dedicated", "custom", P.ProfileName.DEDICATED)],
           "custom": view(P.build_custom_profile(cpu_cores_limit=3, upload_limit_mbps=9.5, not_a_field=1, llm_enabled=False))}
    try:
        P.get_profile("turbo")
        out["unknown"] = "no error"
    except ValueError as exc:
        out["unknown"] = str(exc)
    return out


# (p2p.index_submit is driven in the twenty-third batch, once baseline/shims/trafilatura lets the reference module import)


SCENARIOS.update({f.__name__: f for f in (robots_policy, resource_profiles)})


# ----------------------------------------------------------------------------- thirteenth batch: cross-node credit sync, LAN discovery frames
def credit_sync_across_nodes(pkg, tmp):
    Sy = _m(pkg, "credits.sync")
    L = _m(pkg, "credits.ledger")
    clock = [1_700_000_000.0]
    with mock.patch(f"{pkg}.credits.sync.time.time", side_effect=lambda: clock[0]):
        ledger = L.CreditLedger(tmp / f"ledger-{pkg}.db")
        acts = list(L.ActionType)
        for a, q in ((acts[0], 10), (acts[1], 3), (acts[0], 5)):
            ledger.record_action(a, quantity=q)
        ledger.spend(1.0, reason="search")
        store = Sy.CreditSyncStore(tmp / f"sync-{pkg}.db")
        mgr = Sy.CreditSyncManager(ledger, store, " Owner@Example.ORG ", key_pair=None, local_peer_id="node-local")
        anon = Sy.CreditSyncManager(ledger, Sy.CreditSyncStore(tmp / f"sync-anon-{pkg}.db"), "", local_peer_id="node-x")
        mine = mgr.build_summary()
        same = Sy.CreditSyncManager(ledger, Sy.CreditSyncStore(tmp / f"sync-b-{pkg}.db"), "owner@example.org", local_peer_id="node-b")
        out = {"hash_len": len(mgr.owner_email_hash), "hash_normalised": mgr.owner_email_hash == same.owner_email_hash, "identity": (mgr.has_identity, anon.has_identity),
               "summary": (mine.peer_id, round(mine.total_earned, 4), round(mine.total_spent, 4), round(mine.contribution_score, 4), mine.entry_count, mine.tier, mine.timestamp,
                           mine.owner_email_hash == mgr.owner_email_hash)}

        def remote(peer, earned, spent, *, owner=None, ts=None):
            return Sy.CreditSummary(peer_id=peer, owner_email_hash=mgr.owner_email_hash if owner is None else owner, total_earned=earned, total_spent=spent,
                                    contribution_score=earned, entry_count=3, tier="Tier 1", timestamp=clock[0] if ts is None else ts)

        out["receive"] = [mgr.receive_summary(remote("node-b", 40.0, 4.0), verify_signature=False), mgr.receive_summary(remote("node-c", 10.0, 0.0), verify_signature=False),
                          mgr.receive_summary(remote("node-b", 50.0, 5.0), verify_signature=False),                  # newer summary of a known peer
                          mgr.receive_summary(remote("node-local", 9.0, 0.0), verify_signature=False),               # our own echo
                          mgr.receive_summary(remote("node-d", 9.0, 0.0, owner="someone-else"), verify_signature=False),
                          mgr.receive_summary(remote("node-e", 9.0, 0.0, ts=clock[0] + 4000), verify_signature=False),  # from the future
                          anon.receive_summary(remote("node-b", 1.0, 0.0, owner=""), verify_signature=False)]
        agg = mgr.aggregated_stats()
        out["aggregate"] = (round(agg.total_earned, 4), round(agg.total_spent, 4), round(agg.balance, 4), round(agg.contribution_score, 4), agg.node_count,
                            sorted(s.peer_id for s in agg.peer_summaries))
        out["store"] = (store.peer_count(mgr.owner_email_hash), store.has_peer("node-b"), store.has_peer("node-z"))
        mgr.register_same_owner_peer("node-b")
        out["sync_due"] = [mgr.needs_sync("node-b"), mgr.needs_sync("node-new")]
        clock[0] += 400.0
        out["sync_due_later"] = mgr.needs_sync("node-b")
        out["same_owner"] = mgr.get_same_owner_peers()
        clock[0] += 73 * 3600.0
        out["purged"] = mgr.purge_stale()
        out["after_purge"] = (store.peer_count(mgr.owner_email_hash), mgr.aggregated_stats().node_count)
        wire = mine.to_dict()
        back = Sy.CreditSummary.from_dict(dict(wire, total_earned="lots", entry_count=None, extra="ignored"))
        # (this package's summaries also carry the sender's `public_key` so a receiver can check the signature: a superset)
        out["wire"] = (sorted(k for k in wire if k != "public_key"), back.total_earned, back.entry_count, back.peer_id, Sy.CreditSummary.from_dict({}).tier)
        out["limits"] = (Sy.SUMMARY_TTL_HOURS, Sy.SYNC_INTERVAL_SECONDS, Sy.MAX_PEER_SUMMARIES)
        for m in (mgr, anon, same):
            m.close()
        ledger.close()
    return out


def lan_discovery_frames(pkg, tmp):
    M = _m(pkg, "p2p.mdns")
    import msgpack

    a, b = M.MDNSDiscovery("peer-a", port=4001), M.MDNSDiscovery("peer-b", port=4555)
    with mock.patch(f"{pkg}.p2p.mdns.time.time", return_value=1234.5):
        frame = a._build_announce()
    seen = b._parse_announce(frame, ("192.168.1.20", 5353))
    magic = frame[:len(M.MAGIC)]
    bad = [b._parse_announce(x, ("10.0.0.1", 1)) for x in (b"", b"xx", M.MAGIC, M.MAGIC + b"\xff\xff\xff\xff", b"WRONGMAGIC" + frame[len(M.MAGIC):],
                                                            M.MAGIC + msgpack.packb({"peer_id": "", "port": 1}), M.MAGIC + msgpack.packb({"peer_id": "p", "port": 0}))]
    # (a msgpack frame that is not a map makes the reference raise AttributeError inside its listener thread; here it is dropped --
    #  hardening, asserted for this package alone)
    if pkg == "infomesh_b200":
        assert b._parse_announce(M.MAGIC + msgpack.packb(["not", "a", "map"]), ("10.0.0.1", 1)) is None
    return {"magic": magic.hex(), "payload": {k: v for k, v in sorted(msgpack.unpackb(frame[len(M.MAGIC):], raw=False).items())},
            "seen": (seen.peer_id, seen.host, seen.port), "own_echo": a._parse_announce(frame, ("192.168.1.10", 5353)), "bad": bad,
            "counts": (a.peer_count, dict(a.discovered_peers))}


SCENARIOS.update({f.__name__: f for f in (credit_sync_across_nodes, lan_discovery_frames)})


# ----------------------------------------------------------------------------- fourteenth batch: dashboard data cache, starter snapshot bookkeeping
def dashboard_data_cache(pkg, tmp):
    DC = _m(pkg, "dashboard.data_cache")
    C = _m(pkg, "config")
    LS = _m(pkg, "index.local_store").LocalStore
    import dataclasses
    import sqlite3

    data = tmp / f"data-{pkg}"
    base = C.load_config(tmp / "none.toml")
    cfg = dataclasses.replace(base, node=dataclasses.replace(base.node, data_dir=data), index=dataclasses.replace(base.index, db_path=data / "index.db"))
    cache = DC.DashboardDataCache(cfg, ttl=0.0)
    before = cache.get_stats()                       # no database yet: an empty record, no exception
    out = {"before": (before.document_count, before.top_domains, before.pages_last_hour, before.domain_count, before.recent_docs)}
    data.mkdir(parents=True, exist_ok=True)
    store = LS(data / "index.db")
    hosts = ["docs.python.org", "docs.python.org", "docs.python.org", "en.wikipedia.org", "en.wikipedia.org", "example.org", "sub.example.org", "example.org"]
    for i, h in enumerate(hosts):
        store.add_document(url=f"https://{h}/page/{i}", title=f"Title {i}" if i % 3 else "", text=f"body text number {i} " * 10, raw_html_hash=f"r{i}", text_hash=f"t{i}")
    store.close()
    conn = sqlite3.connect(data / "index.db")        # spread the crawl times: three within the last hour, the rest older
    import time as _t

    now = _t.time()
    for i in range(len(hosts)):
        conn.execute("UPDATE documents SET crawled_at = ? WHERE url = ?", (now - (600 * i if i < 3 else 7200 + 100 * i), f"https://{hosts[i]}/page/{i}"))
    conn.commit()
    conn.close()
    st = cache.get_stats()
    out["after"] = {"docs": st.document_count, "top": [tuple(x) for x in st.top_domains][:4], "last_hour": st.pages_last_hour, "domains": st.domain_count,
                    "recent": [(d.url, d.title) for d in st.recent_docs], "last_crawl_recent": now - st.last_crawl_at < 5}
    cache.set_ttl(1000.0)
    st2 = cache.get_stats()
    out["cached_same_object"] = st2 is cache.get_stats()
    cache.close()
    cache.close()                                    # idempotent
    return out


def starter_snapshot_bookkeeping(pkg, tmp):
    S = _m(pkg, "index.starter")
    info = S.StarterAssetInfo("https://github.com/x/y/releases/download/v1/starter.infomesh-snapshot", 5 * 1024 * 1024 + 512, "v1.2.3", "2024-05-06T07:08:09Z")
    d = tmp / f"starter-{pkg}"
    d.mkdir()
    out = {"needs": [S.needs_starter(n) for n in (0, 9, 10, 11, 10_000)], "size_mb": round(info.size_mb, 4), "empty_cache": S._read_cache(d)}
    with mock.patch(f"{pkg}.index.starter.time.time", return_value=1_000_000.0):
        S._write_cache(d, info)
        back = S._read_cache(d)
    out["round_trip"] = (back.download_url, back.size_bytes, back.release_tag, back.created_at)
    with mock.patch(f"{pkg}.index.starter.time.time", return_value=1_000_000.0 + 30 * 86400):
        out["expired"] = S._read_cache(d)
    for f in d.iterdir():
        f.write_text("{not json")
    out["corrupt"] = S._read_cache(d)
    return out


SCENARIOS.update({f.__name__: f for f in (dashboard_data_cache, starter_snapshot_bookkeeping)})


# ----------------------------------------------------------------------------- fifteenth batch: the Python SDK over a prepared data directory
SDK_DOCS = [("https://docs.python.org/3/library/asyncio.html", "asyncio — Asynchronous I/O", "asyncio is a library to write concurrent code using the async await syntax. " * 6, "en"),
            ("https://realpython.com/async-io-python/", "Async IO in Python: A Complete Walkthrough", "This tutorial covers asyncio coroutines event loop and tasks in python. " * 6, "en"),
            ("https://doc.rust-lang.org/book/ch04-01-what-is-ownership.html", "What is Ownership?", "Ownership is a set of rules that govern how a rust program manages memory. " * 6, "en"),
            ("https://ko.wikipedia.org/wiki/파이썬", "파이썬", "파이썬은 고급 프로그래밍 언어로 플랫폼에 독립적이며 인터프리터식 객체지향적 동적 타이핑 대화형 언어이다. " * 6, "ko"),
            ("https://blog.example.org/python-packaging", "Packaging python projects", "How to package python projects with pyproject toml and publish them to an index. " * 6, "en")]


def sdk_client(pkg, tmp):
    K = _m(pkg, "sdk.client")
    LS = _m(pkg, "index.local_store").LocalStore
    data = tmp / f"sdk-{pkg}"
    data.mkdir()
    store = LS(str(data / "index.db"))
    for i, (url, title, text, lang) in enumerate(SDK_DOCS):
        store.add_document(url=url, title=title, text=text, raw_html_hash=f"r{i}", text_hash=f"t{i}", language=lang)
    store.close()
    try:
        client = K.InfoMeshClient(data_dir=str(data), gpu=False)        # this package can attach a GPU index; the CPU contract is compared
    except TypeError:
        client = K.InfoMeshClient(data_dir=str(data))
    with client as c:
        def urls(**kw):
            q = kw.pop("q")
            return [r.url for r in c.search(q, **kw)]

        out = {"python": urls(q="python"), "asyncio_top": urls(q="asyncio event loop", limit=1), "offset": urls(q="python", limit=1, offset=1), "none": urls(q="kubernetes"),
               "lang": urls(q="python", language="en"), "include": urls(q="python", include_domains=["realpython.com"]),
               "exclude": urls(q="python", exclude_domains=["docs.python.org", "realpython.com"]), "operators": urls(q='python AND "asyncio" NOT (rust) *'),
               "korean": urls(q="파이썬"), "shape": sorted(c.search("python", limit=1)[0].to_dict()),
               "scores_sorted": (lambda s: s == sorted(s, reverse=True))([r.score for r in c.search("python")]),
               "suggest": [c.suggest(p, limit=3) for p in ("pyt", "async", "zzz", "")], "stats_docs": c.get_stats().get("document_count")}
    return out


SCENARIOS.update({f.__name__: f for f in (sdk_client,)})


# ----------------------------------------------------------------------------- sixteenth batch: replica placement (same peers chosen by both implementations)
def replica_placement(pkg, tmp):
    R = _m(pkg, "p2p.replication")
    peers = [f"12D3KooW{name}" for name in ("Alpha", "Bravo", "Charlie", "Delta", "Echo", "Foxtrot", "Golf", "Hotel")]
    me = peers[3]
    urls = [f"https://e.com/page/{i}" for i in range(12)] + ["https://한국어.example/문서", ""]

    class Host:
        def get_connected_peers(self):
            return list(peers)

    async def reference_way(factor):
        rep = R.Replicator(Host(), None, me, replication_factor=factor)
        return [await rep._find_replica_peers(u) for u in urls]

    def this_way(factor):
        return [R.replica_peers(u, [p for p in peers if p != me], factor) for u in urls]

    pick = (lambda f: asyncio.run(reference_way(f))) if pkg == "infomesh" else this_way
    return {"three": pick(3), "one": pick(1), "more_than_available": pick(20), "default_factor": R.DEFAULT_REPLICATION_FACTOR}


SCENARIOS.update({f.__name__: f for f in (replica_placement,)})


# ----------------------------------------------------------------------------- seventeenth batch: the services layer (needs baseline/shims/trafilatura)
PAGES = [("https://docs.example.org/asyncio", "asyncio guide", "asyncio lets you write concurrent code with async and await. " * 8, "en"),
         ("https://docs.example.org/asyncio?utm_source=x", "asyncio guide (tracked link)", "asyncio lets you write concurrent code with async and await. " * 8, "en"),
         ("https://blog.example.org/rust", "Rust ownership", "ownership and borrowing make rust memory safe without a garbage collector. " * 8, None),
         ("https://ko.example.org/python", "파이썬", "파이썬은 배우기 쉬운 프로그래밍 언어입니다. " * 12, "ko")]


class _RecordingVectors:
    def __init__(self):
        self.added = []

    def add_document(self, **kw):
        self.added.append((kw["doc_id"], kw["url"], kw.get("language")))

    def search(self, query, limit=10):
        return []


def services_indexing(pkg, tmp):
    S = _m(pkg, "services")
    PP = _m(pkg, "crawler.parser").ParsedPage
    LS = _m(pkg, "index.local_store").LocalStore
    H = _m(pkg, "hashing")
    store, vec = LS(tmp / f"svc-{pkg}.db"), _RecordingVectors()
    ids = []
    for url, title, text, lang in PAGES:
        page = PP(url=url, title=title, text=text, language=lang, raw_html_hash=H.content_hash("<html>" + text), text_hash=H.content_hash(text))
        ids.append(S.index_document(page, store, vec, js_required=url.endswith("rust")))
    again = S.index_document(PP(url=PAGES[0][0], title="changed title", text=PAGES[0][2], language="en", raw_html_hash="x", text_hash=H.content_hash(PAGES[0][2])), store, None)
    out = {"ids": [i is not None for i in ids], "distinct": len({i for i in ids if i is not None}), "again": again is not None, "vectors": [(u, lg) for _, u, lg in vec.added],
           "count": store.get_stats()["document_count"], "js": [bool(getattr(store.get_document_by_url(u), "js_required", False)) if store.get_document_by_url(u) else None
                                                                  for u, *_ in PAGES],
           "langs": [getattr(store.get_document_by_url(u), "language", None) if store.get_document_by_url(u) else None for u, *_ in PAGES],
           "found": [r.url for r in _m(pkg, "search.query").search_local(store, "asyncio await", limit=5).results]}
    store.close()
    return out


def parser_bookkeeping(pkg, tmp):
    """Through ``extract_content``: only the fields that do not depend on extraction quality (hashes, language attribute, the
    too-short rule) -- the reference runs on the tag-stripping stand-in for trafilatura here."""
    P = _m(pkg, "crawler.parser")
    H = _m(pkg, "hashing")
    body = "<p>" + "Tensor memory holds the accumulators of the matrix units. " * 10 + "</p>"
    html = f'<html lang="en-GB"><head><title>  Tensor   memory </title><style>p{{color:red}}</style></head><body>{body}<script>var x = 1;</script></body></html>'
    page = P.extract_content(html, "https://e.com/tmem")
    short = P.extract_content("<html><body><p>too short</p></body></html>", "https://e.com/short")
    empty = P.extract_content("", "https://e.com/empty")
    given = P.extract_content(html, "https://e.com/tmem", raw_hash="abc123")
    return {"url": page.url, "title": " ".join(page.title.split()), "language": page.language, "raw_hash_is_html_hash": page.raw_html_hash == H.content_hash(html),
            "text_hash_is_text_hash": page.text_hash == H.content_hash(page.text), "script_dropped": "var x" not in page.text, "style_dropped": "color:red" not in page.text,
            "has_body": "Tensor memory holds" in page.text, "short": short is None, "empty": empty is None, "raw_hash_passed_through": given.raw_html_hash}


SCENARIOS.update({f.__name__: f for f in (services_indexing, parser_bookkeeping)})


# ----------------------------------------------------------------------------- eighteenth batch: the crawl worker over a scripted web site
def _site():
    long_a = "<p>" + "Tensor memory holds accumulators while the matrix units run. " * 12 + "</p>"
    long_c = "<p>" + "Kademlia keeps buckets of contacts ordered by the XOR distance to the local id. " * 12 + "</p>"
    return {
        "https://site.example/robots.txt": (200, "User-agent: *\nDisallow: /blocked\nCrawl-delay: 0\nSitemap: https://site.example/sitemap.xml\n", "text/plain"),
        "https://site.example/a": (200, f'<html lang="en"><head><title>Page A</title><link rel="alternate" type="application/rss+xml" href="/feed.xml"></head><body>{long_a}'
                                        '<a href="/b">b</a> <a href="https://site.example/c#frag">c</a> <a href="https://other.example/x">ext</a> <a href="mailto:x@y.z">m</a>'
                                        '<a href="/a">self</a> <a href="javascript:void(0)">js</a> <a href="/file.pdf">pdf</a></body></html>', "text/html"),
        "https://site.example/b": (200, f"<html><head><title>Page B</title></head><body>{long_a}</body></html>", "text/html"),            # same text as /a
        "https://site.example/c": (200, f"<html><head><title>Page C</title></head><body>{long_c}</body></html>", "text/html"),
        "https://site.example/short": (200, "<html><body><p>tiny</p></body></html>", "text/html"),
        "https://site.example/boom": (500, "server error", "text/html"),
        "https://site.example/gone": (404, "not found", "text/html"),
        "https://site.example/blocked/secret": (200, f"<html><body>{long_c}</body></html>", "text/html"),
        "https://site.example/image.png": (200, "\x89PNG....", "image/png"),
    }


class _SiteClient:
    """httpx.AsyncClient stand-in: ``get`` answers from the table and counts calls."""
    is_closed = False

    def __init__(self):
        self.site, self.calls = _site(), []

    async def get(self, url, **kw):
        import httpx

        url = str(url)
        self.calls.append(url)
        status, body, ctype = self.site.get(url, (404, "", "text/html"))
        req = httpx.Request("GET", url)
        return httpx.Response(status, text=body, headers={"content-type": ctype}, request=req)

    async def aclose(self):
        self.is_closed = True


def crawl_worker_over_a_scripted_site(pkg, tmp):
    W = _m(pkg, "crawler.worker")
    Sch = _m(pkg, "crawler.scheduler").Scheduler
    DD = _m(pkg, "crawler.dedup").DeduplicatorDB
    RC = _m(pkg, "crawler.robots").RobotsChecker
    C = _m(pkg, "config")
    import dataclasses
    import socket

    crawl_cfg = dataclasses.replace(C.load_config(tmp / "none.toml").crawl, politeness_delay=0.1, max_depth=2, urls_per_hour=10000, pending_per_domain=100)

    def public_dns(host, *a, **k):                      # every name resolves to a public address; literal IPs keep their own
        try:
            import ipaddress

            ip = str(ipaddress.ip_address(host))
        except ValueError:
            ip = "93.184.216.34"
        return [(socket.AF_INET, socket.SOCK_STREAM, 6, "", (ip, 0))]

    async def no_sleep(_s):
        return None

    async def go():
        sched, dedup = Sch(politeness_delay=0.0, urls_per_hour=0, pending_per_domain=100, max_depth=2), DD(str(tmp / f"w-dedup-{pkg}.db"))
        worker = W.CrawlWorker(crawl_cfg, sched, dedup, RC(crawl_cfg.user_agent))
        client = _SiteClient()
        worker._client = client
        rows = []
        with mock.patch("socket.getaddrinfo", public_dns), mock.patch(f"{pkg}.crawler.worker.asyncio.sleep", no_sleep):
            for url, kw in (("https://site.example/a", {}), ("https://site.example/b", {}), ("https://site.example/c", {"depth": 1}), ("https://site.example/short", {}),
                            ("https://site.example/boom", {}), ("https://site.example/gone", {}), ("https://site.example/blocked/secret", {}), ("https://site.example/image.png", {}),
                            ("http://127.0.0.1/admin", {}), ("http://169.254.169.254/latest/meta-data", {}), ("ftp://site.example/x", {}), ("https://site.example/a", {}),
                            ("https://site.example/a", {"force": True}), ("https://site.example/a?utm_source=nl", {})):
                r = await worker.crawl_url(url, **kw)
                rows.append((url, kw.get("force", False), r.success, (r.error or "").split(":")[0][:40], r.page.title if r.page else None,
                             sorted(l for l in r.discovered_links if "site.example" in l or "other.example" in l), sorted(getattr(r, "discovered_feeds", []) or []), r.js_required))
        await worker.close()
        out = {"rows": rows, "boom_attempts": client.calls.count("https://site.example/boom"), "robots_fetches": client.calls.count("https://site.example/robots.txt"),
               "seen": [dedup.is_url_seen(u) for u in ("https://site.example/a", "https://site.example/b", "https://site.example/c", "https://site.example/short")],
               "scheduled": sched.pending_count}
        dedup.close()
        return out

    return asyncio.run(go())


SCENARIOS.update({f.__name__: f for f in (crawl_worker_over_a_scripted_site,)})


# ----------------------------------------------------------------------------- nineteenth batch: what an AppContext wires up for each role
def app_context_wiring(pkg, tmp):
    S = _m(pkg, "services")
    C = _m(pkg, "config")
    import dataclasses

    names = ("store", "vector_store", "key_pair", "dedup", "robots", "scheduler", "worker", "link_graph", "ledger", "governor", "feed_monitor", "priority_queue", "llm_backend",
             "index_submit_sender", "index_submit_receiver", "distributed_index", "p2p_node", "credit_sync_manager")
    out = {}
    for role in ("full", "crawler", "search"):
        base = C.load_config(tmp / "none.toml")
        data = tmp / f"ctx-{pkg}-{role}"
        cfg = dataclasses.replace(base, node=dataclasses.replace(base.node, data_dir=data, role=role), index=dataclasses.replace(base.index, db_path=data / "index.db", vector_search=False),
                                  network=dataclasses.replace(base.network, bootstrap_dns=False, bootstrap_github=False))
        ctx = S.AppContext(cfg)
        try:
            out[role] = {n: getattr(ctx, n, None) is not None for n in names}
            out[role]["keys_on_disk"] = (data / "keys" / "private.pem").exists()
            out[role]["files"] = sorted(p.name for p in data.iterdir() if p.suffix == ".db")
            out[role]["docs"] = ctx.store.get_stats()["document_count"]
            out[role]["peer_id_len"] = len(ctx.key_pair.peer_id) if getattr(ctx, "key_pair", None) else 0
        finally:
            closer = getattr(ctx, "close_async", None)
            asyncio.run(closer()) if closer else ctx.close()
    return out


SCENARIOS.update({f.__name__: f for f in (app_context_wiring,)})


# ----------------------------------------------------------------------------- twentieth batch: publishing and page fetching through the services layer
class _RecordingNetwork:
    """Either a p2p node (``publish_document_to_network`` / ``publish_batch_to_network``) or a distributed index (``publish_document`` /
    ``publish_batch``), recording what it is asked to publish."""

    def __init__(self, kind: str, fail_on: int | None = None):
        self.kind, self.calls, self.fail_on = kind, [], fail_on
        if kind == "node":
            self.publish_document_to_network = self._one_positional
            self.publish_batch_to_network = self._batch
        else:
            self.publish_document = self._one_keyword
            self.publish_batch = self._batch

    async def _one_positional(self, doc_id, url, title, text):
        self.calls.append(("doc", doc_id, url))
        return 7

    async def _one_keyword(self, *, doc_id, url, title, text):
        self.calls.append(("doc", doc_id, url))
        return 5

    async def _batch(self, documents):
        self.calls.append(("batch", len(documents)))
        if self.fail_on is not None and len(self.calls) == self.fail_on:
            raise ConnectionError("dht unreachable")
        return 10 * len(documents)


def services_publish_and_fetch(pkg, tmp):
    S = _m(pkg, "services")
    PP = _m(pkg, "crawler.parser").ParsedPage
    LS = _m(pkg, "index.local_store").LocalStore
    store = LS(tmp / f"pub-{pkg}.db")
    for i in range(7):
        store.add_document(url=f"https://e.com/{i}", title=f"T{i}", text=f"document number {i} about kernels and tiles " * 8, raw_html_hash=f"r{i}", text_hash=f"t{i}")
    store.add_document(url="https://e.com/paywall", title="Premium", text="Subscribe to continue reading this article. " * 6, raw_html_hash="rp", text_hash="tp")
    page = PP(url="https://e.com/0", title="T0", text="x", language="en", raw_html_hash="r", text_hash="t")

    async def go():
        node, index, broken = _RecordingNetwork("node"), _RecordingNetwork("index"), _RecordingNetwork("index", fail_on=2)
        single = [await S.publish_document_to_network(page, 3, p2p_node=node), await S.publish_document_to_network(page, 3, distributed_index=index),
                  await S.publish_document_to_network(page, None, p2p_node=node), await S.publish_document_to_network(page, 3),
                  await S.publish_document_to_network(page, 3, p2p_node=node, distributed_index=index)]
        rn, ri = _RecordingNetwork("node"), _RecordingNetwork("index")
        bulk = [await S.republish_local_index(store, p2p_node=rn, batch_size=3), await S.republish_local_index(store, distributed_index=ri, batch_size=5, limit=6),
                await S.republish_local_index(store), await S.republish_local_index(store, distributed_index=broken, batch_size=3),
                await S.republish_local_index(store, distributed_index=_RecordingNetwork("index"), batch_size=0)]
        return single, (node.calls, index.calls), bulk, (rn.calls, ri.calls, broken.calls)

    single, single_calls, bulk, bulk_calls = asyncio.run(go())
    old = store.get_document_by_url("https://e.com/1").crawled_at

    def view(r):
        return (r.success, r.url, r.title, len(r.text), r.is_cached, r.is_stale, r.is_paywall, (r.error or "").split(":")[0])

    with mock.patch(f"{pkg}.services.time.time", return_value=old + 10):
        fresh = S.fetch_page("https://e.com/1", store=store, worker=None)
        clipped = S.fetch_page("https://e.com/1", store=store, worker=None, max_size_bytes=50)
    with mock.patch(f"{pkg}.services.time.time", return_value=old + 8 * 86400):
        stale = S.fetch_page("https://e.com/1", store=store, worker=None)
    out = {"single": single, "single_calls": single_calls, "bulk": bulk, "bulk_calls": bulk_calls,
           "fetch": [view(fresh), view(clipped), view(stale), view(S.fetch_page("https://e.com/unknown", store=store, worker=None)),
                     view(S.fetch_page("http://127.0.0.1/x", store=store, worker=None)), view(S.fetch_page("file:///etc/passwd", store=store, worker=None))],
           "paywall": [S.is_paywall_content(t) for t in ("Subscribe to continue reading", "plain article text about kernels", "", "Sign in to read the full story. Already a subscriber?")],
           "truncate": [S._truncate_to_bytes(t, n) for t, n in (("hello world", 5), ("héllo", 2), ("한국어 텍스트", 7), ("short", 100), ("", 3))]}
    store.close()
    return out


SCENARIOS.update({f.__name__: f for f in (services_publish_and_fetch,)})


# ----------------------------------------------------------------------------- twenty-first batch: crawl -> link graph -> index -> publish, and the peer-facing search function
def crawl_and_index_pipeline(pkg, tmp):
    S = _m(pkg, "services")
    W = _m(pkg, "crawler.worker")
    Sch = _m(pkg, "crawler.scheduler").Scheduler
    DD = _m(pkg, "crawler.dedup").DeduplicatorDB
    RC = _m(pkg, "crawler.robots").RobotsChecker
    LS = _m(pkg, "index.local_store").LocalStore
    LG = _m(pkg, "index.link_graph").LinkGraph
    C = _m(pkg, "config")
    import dataclasses
    import socket

    base = C.load_config(tmp / "none.toml")
    data = tmp / f"pipe-{pkg}"
    cfg = dataclasses.replace(base, node=dataclasses.replace(base.node, data_dir=data), index=dataclasses.replace(base.index, db_path=data / "index.db", vector_search=False))
    data.mkdir()

    def public_dns(host, *a, **k):
        return [(socket.AF_INET, socket.SOCK_STREAM, 6, "", ("93.184.216.34", 0))]

    async def no_sleep(_s):
        return None

    async def go():
        store, dedup, graph = LS(data / "index.db"), DD(str(data / "dedup.db")), LG(str(data / "links.db"))
        worker = W.CrawlWorker(cfg.crawl, Sch(politeness_delay=0.0, urls_per_hour=0, pending_per_domain=100, max_depth=2), dedup, RC(cfg.crawl.user_agent))
        worker._client = _SiteClient()
        net, vec = _RecordingNetwork("node"), _RecordingVectors()
        rows = []
        with mock.patch("socket.getaddrinfo", public_dns), mock.patch(f"{pkg}.crawler.worker.asyncio.sleep", no_sleep):
            for url, kw in (("https://site.example/a", {}), ("https://site.example/c", {}), ("https://site.example/gone", {}), ("https://site.example/a", {}),
                            ("https://site.example/a", {"force": True})):
                r = await S.crawl_and_index(url, worker=worker, store=store, vector_store=vec, p2p_node=net, link_graph=graph, **kw)
                rows.append((r.success, r.url, r.title, r.text_length > 0, r.links_discovered, (r.error or "").split(":")[0]))
        await worker.close()
        out = {"rows": rows, "docs": store.get_stats()["document_count"], "published": net.calls, "vectors": len(vec.added),
               "out_links": sorted(graph.get_outlinks("https://site.example/a")) if hasattr(graph, "get_outlinks") else None}
        search = S.create_local_search_fn(cfg)
        hits = await search("tensor memory accumulators", 3)
        out["peer_search"] = [(h["url"], h["title"], sorted(h)) for h in hits]
        out["peer_search_empty"] = await search("zzzqqq", 3)
        store.close(), dedup.close(), graph.close()
        return out

    return asyncio.run(go())


SCENARIOS.update({f.__name__: f for f in (crawl_and_index_pipeline,)})


# ----------------------------------------------------------------------------- twenty-second batch: coercion of untrusted peer payloads in the query router
def router_payload_coercion(pkg, tmp):
    R = _m(pkg, "p2p.routing")
    odd = [None, True, False, 0, 7, -3, 2.9, -2.9, float("nan"), float("inf"), -float("inf"), "12", " 12 ", "1e3", "0x10", "abc", "", b"5", [1], {"a": 1}, (2,), 10 ** 30, "9" * 400]
    return {"str": [R._payload_str(v) for v in odd] + [R._payload_str(None, default="d"), R._payload_str(5, default="d")],
            "int": [R._payload_int(v) for v in odd] + [R._payload_int("x", default=9), R._payload_int(None, default=-1)],
            "float": [repr(R._payload_float(v)) for v in odd] + [repr(R._payload_float("x", default=0.5))]}


SCENARIOS.update({f.__name__: f for f in (router_payload_coercion,)})


# ----------------------------------------------------------------------------- twenty-third batch: crawler -> indexer submissions (enterprise split)
def index_submission_flow(pkg, tmp):
    S = _m(pkg, "p2p.index_submit")
    P = _m(pkg, "p2p.protocol")
    C = _m(pkg, "config")
    PP = _m(pkg, "crawler.parser").ParsedPage
    LS = _m(pkg, "index.local_store").LocalStore
    H = _m(pkg, "hashing")
    import dataclasses

    base = C.load_config(tmp / "none.toml")

    def cfg(**net):
        return dataclasses.replace(base, network=dataclasses.replace(base.network, **net))

    body = "Document body about tensors and kernels on one node. " * 12
    page = PP(url="https://e.com/doc", title="Doc", text=body, language="en", raw_html_hash=H.content_hash("<html>" + body), text_hash=H.content_hash(body))
    sender = S.IndexSubmitSender(cfg(index_submit_peers=["/ip4/10.0.0.9/tcp/4001/p2p/12D3KooWIndexer"]))
    frame = sender.build_submit_message(page, ["https://e.com/next", "https://e.com/other"])
    kind, payload = P.decode_message(frame)
    sender.record_sent()
    sender.record_sent()
    sender.record_error()
    out = {"kind": int(kind), "payload_keys": sorted(payload), "payload": {k: payload[k] for k in ("url", "title", "language", "text_hash", "raw_html_hash", "peer_id", "discovered_links")},
           "signature_empty_without_key": payload["signature"] in (b"", ""), "sender": {"peers": sender.submit_peers, "stats": dict(sorted(sender.stats.items()))}}
    store = LS(tmp / f"submit-{pkg}.db")
    open_recv = S.IndexSubmitReceiver(cfg(), store)

    def ack_view(a):
        return (a.url, a.success, bool(a.doc_id) if hasattr(a, "doc_id") else None, (a.error or "")[:30])

    acks = [open_recv.handle_submit(dict(payload)), open_recv.handle_submit(dict(payload)),                                   # second: duplicate content
            open_recv.handle_submit(dict(payload, url="https://e.com/two", text=body + " more", text_hash=H.content_hash(body + " more")))]
    out["open"] = {"acks": [ack_view(a) for a in acks], "stats": dict(sorted(open_recv.stats.items())), "docs": store.get_stats()["document_count"],
                   "allowed_anyone": [open_recv.is_peer_allowed(p) for p in ("x", "")]}
    ack_kind, ack_payload = P.decode_message(open_recv.build_ack_message(acks[0]))
    out["ack_wire"] = (int(ack_kind), sorted(ack_payload))
    gated = S.IndexSubmitReceiver(cfg(peer_acl=["12D3KooWCrawler"]), store)
    out["gated"] = {"allowed": [gated.is_peer_allowed(p) for p in ("12D3KooWCrawler", "12D3KooWStranger", "")],
                    "stranger": ack_view(gated.handle_submit(dict(payload, url="https://e.com/three", peer_id="12D3KooWStranger"))),
                    "stats": dict(sorted(gated.stats.items()))}
    store.close()
    return out


SCENARIOS.update({f.__name__: f for f in (index_submission_flow,)})
