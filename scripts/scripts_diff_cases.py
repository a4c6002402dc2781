"""Inputs for ``diff_vs_reference.py behaviour``: (module, function, [(args, kwargs), ...]).  Pure functions only."""
T1 = ("Blackwell keeps MMA accumulators in tensor memory. Tensor memory is 256 KB per SM and is read back with tcgen05.ld. "
      "The fifth generation tensor cores are issued by a single elected thread.\n\nTMA moves tiles between global and shared "
      "memory. It supports multicast across a cluster! Does it need descriptors? Yes: a CUtensorMap per tensor.")
T2 = "파이썬은 배우기 쉬운 프로그래밍 언어입니다. 데이터 분석과 웹 개발에 널리 사용됩니다. 많은 개발자들이 파이썬을 좋아합니다."
T3 = "東京は日本の首都です。人口は約1400万人で、世界最大級の都市圏を形成しています。"
T4 = "Der schnelle braune Fuchs springt über den faulen Hund und die Katze schläft auf dem Sofa in der Sonne."
HTML = ('<html lang="en"><head><title>Doc</title><link rel="alternate" type="application/rss+xml" href="/feed.xml">'
        '<link rel="canonical" href="https://example.org/canon"><script type="application/ld+json">{"@type":"Article","headline":"H"}</script>'
        '<meta property="og:title" content="OG T"></head><body><div id="root"></div><a href="/a">a</a><a href="https://o.org/b#f">b</a>'
        '<pre><code class="language-python">print(1)</code></pre><table><tr><th>k</th></tr><tr><td>v</td></tr></table></body></html>')
RSS = ('<?xml version="1.0"?><rss><channel><title>F</title><item><title>A</title><link>https://e.org/a</link><pubDate>Mon, 01 Jan 2024 00:00:00 GMT</pubDate>'
       '</item><item><title>B</title><link>https://e.org/b</link></item></channel></rss>')


def _c(*args, **kw):
    return (args, kw)


CASES = [
    ("hashing", "content_hash", [_c("abc"), _c(""), _c("한국어")]),
    ("hashing", "short_hash", [_c("abc"), _c("abc", 8)]),
    ("crawler.simhash", "simhash", [_c(T1), _c("a b"), _c(""), _c(T2), _c("one two three four five")]),
    ("crawler.simhash", "hamming_distance", [_c(0, 0), _c(0xFF, 0x0F), _c(2 ** 64 - 1, 0)]),
    ("crawler.dedup", "normalize_url", [_c("HTTPS://Example.org:443/a/b/?b=2&a=1&utm_source=x#frag"), _c("http://example.org:80/"), _c("https://e.org/a//b/./c/../d"),
                                        _c("https://e.org/path/?fbclid=1&q=x"), _c("https://e.org"), _c("https://EXAMPLE.org/A?Z=1&gclid=3")]),
    ("search.passage", "split_passages", [_c(T1), _c(T1, max_length=80), _c("short"), _c(""), _c("word " * 400)]),
    ("search.passage", "select_best_passage", [_c(T1, "tensor memory"), _c(T1, "multicast cluster", max_length=100), _c("", "q"), _c(T1, "")]),
    ("search.passage", "highlight_terms", [_c("Tensor memory holds tensors", "tensor"), _c("abc", ""), _c("a.b (c)", "a.b")]),
    ("search.passage", "title_match_score", [_c("Tensor Memory Guide", "tensor memory"), _c("", "x"), _c("Unrelated", "tensor"), _c("tensor", "tensor memory guide")]),
    ("search.passage", "url_path_score", [_c("https://e.org/docs/tensor-memory/guide.html", "tensor memory"), _c("https://e.org/", "x"), _c("https://tensor.org/a", "tensor")]),
    ("search.passage", "classify_intent", [_c("how to install cuda"), _c("github login"), _c("buy gpu"), _c("what is tmem"), _c("python"), _c("")]),
    ("search.nlp", "remove_stop_words", [_c("the quick brown fox is in the house".split()), _c("der schnelle fuchs und die katze".split(), "de"), _c([]), _c(["the", "the"])]),
    ("search.nlp", "expand_query", [_c("python tutorial"), _c("fast car"), _c("xyzzy"), _c("js error fix")]),
    ("search.nlp", "did_you_mean", [_c("pyhton", ["python", "java", "rust"]), _c("python", ["python"]), _c("zzzzzz", ["python"]), _c("", [])]),
    ("search.nlp", "parse_natural_query", [_c("site:docs.python.org asyncio"), _c("tutorials in korean"), _c("papers since 2023"), _c("plain query")]),
    ("search.cjk", "is_cjk_text", [_c(T2), _c(T3), _c("english"), _c(""), _c("mixed 한글 text")]),
    ("search.cjk", "cjk_bigrams", [_c("東京都"), _c("ab"), _c(""), _c("한국어 검색")]),
    ("search.cjk", "cjk_trigrams", [_c("東京都庁"), _c("한국")]),
    ("search.cjk", "tokenize_query_cjk", [_c("東京 tower"), _c("hello world"), _c("한국어 검색 엔진")]),
    ("search.cjk", "recommend_tokenizer", [_c(T2), _c("english text")]),
    ("search.cjk", "segment_korean", [_c("파이썬은 좋은 언어입니다"), _c("")]),
    ("crawler.lang_detect", "detect_language", [_c(T1), _c(T2), _c(T3), _c(T4), _c(""), _c("12345 !!!"), _c("Привет мир, как дела сегодня")]),
    ("crawler.rss", "parse_feed_xml", [_c(RSS, "https://e.org/feed"), _c("<feed xmlns='http://www.w3.org/2005/Atom'><title>T</title><entry><title>E</title><link href='https://e.org/e'/><updated>2024-01-02T03:04:05Z</updated></entry></feed>", "https://e.org/atom"), _c("garbage", "https://e.org/x")]),
    ("crawler.rss", "discover_feeds", [_c(HTML, "https://example.org/page"), _c("<html></html>", "https://e.org")]),
    ("crawler.structured", "extract_structured_data", [_c(HTML), _c("<html></html>")]),
    ("crawler.js_detect", "detect_js_requirement", [_c(HTML), _c("<html><body><p>" + "text " * 200 + "</p></body></html>"), _c("<html><body><noscript>Please enable JavaScript to run this app</noscript><div id='app'></div><script src='/static/js/main.chunk.js'></script></body></html>"), _c("")]),
    ("crawler.content_extract", "extract_code_blocks", [_c(HTML)]),
    ("crawler.content_extract", "extract_tables", [_c(HTML)]),
    ("crawler.diff", "compute_diff", [_c("a\nb\nc", "a\nx\nc"), _c("", "new"), _c("same", "same")]),
    ("index.distributed", "extract_keywords", [_c(T1), _c("the and of"), _c(""), _c("GPU gpu Gpu cuda", 2)]),
    ("index.ranking", "freshness_score", [_c(0.0, now=86400.0 * 7), _c(100.0, now=50.0), _c(0.0, now=86400.0 * 700)]),
    ("index.ranking", "normalize_bm25", [_c(3.0, max_score=1.0), _c(0.0, max_score=1.0), _c(1.0, max_score=0.0)]),
    ("search.query", "_sanitize_fts_query", [_c('hello "world" (test)'), _c("a AND b OR NOT c NEAR d"), _c("***"), _c("x" * 2000), _c("c++ & c#"), _c("what's new: gpu^2 {x}")]),
    ("search.quality", "ndcg_at_k", [_c([3, 2, 3, 0, 1, 2], 6), _c([], 5), _c([0, 0], 2), _c([1.0, 0.5], 1)]),
    ("search.quality", "mrr", [_c([[0, 1, 0], [1, 0], [0, 0]]), _c([])]),
    ("search.quality", "extract_temporal_hint", [_c("news today"), _c("events last month"), _c("python 2023 release"), _c("plain")]),
    ("search.rag", "extract_entities", [_c("Guido van Rossum created Python in 1991 at CWI in Amsterdam. Contact: g@python.org, see https://python.org")]),
    ("search.rag", "compute_toxicity_score", [_c("you are a stupid idiot"), _c("nice weather today"), _c("")]),
    ("search.rag", "build_summary_prompt", [_c("q", [{"title": "T", "url": "https://u", "snippet": "S"}])]),
    ("data_quality", "extract_citations", [_c("See doi:10.1000/xyz123 and arXiv:2301.12345, RFC 9110, ISBN 978-3-16-148410-0 and https://e.org/x.")]),
    ("data_quality", "compute_trust_grade", [_c(0.95), _c(0.7), _c(0.5), _c(0.2), _c(0.0)]),
    ("data_quality", "compute_freshness_indicator", [_c(1_000_000.0, now=1_000_000.0 + 3600), _c(0.0, now=86400.0 * 400), _c(1_000_000.0, now=1_000_000.0 + 86400 * 10)]),
    ("p2p.protocol", "keyword_to_dht_key", [_c("Rust"), _c("")]),
    ("p2p.protocol", "url_to_dht_key", [_c("https://e.org/")]),
    ("security", "validate_url", [_c("https://example.org/"), _c("ftp://x"), _c("http://127.0.0.1/"), _c("http://169.254.169.254/"), _c("http://localhost/"), _c("https://[::1]/"),
                                  _c("http://10.1.2.3/"), _c("https://user:pw@example.org/"), _c("http://0x7f.1/"), _c("")]),
    ("errors", "format_error", [_c("E001"), _c("NOPE")]),
    ("api.extensions", "get_completion_commands", [_c()]),
    ("p2p.sybil", "compute_pow_hash", [_c(b"k" * 32, 0), _c(b"k" * 32, 12345)]),
    ("p2p.sybil", "derive_node_id", [_c(b"k" * 32, 7)]),
    ("p2p.sybil", "verify_pow", [_c(b"k" * 32, 1, 1), _c(b"k" * 32, 1, 30)]),
    ("credits.scheduling", "is_off_peak_at", [_c(hour=23), _c(hour=3), _c(hour=12), _c(hour=7), _c(hour=22, start=22, end=6), _c(hour=6, start=22, end=6)]),
    ("search.facets", "highlight_snippet", [_c("Tensor memory is fast memory", "memory tensor"), _c("abc", "")]),
    ("search.facets", "compute_facets", [_c(lambda p: _ranked(p))]),
    ("search.facets", "cluster_results", [_c(lambda p: _ranked(p))]),
    ("search.facets", "dedup_results", [_c(lambda p: _ranked(p))]),
    ("search.cross_validate", "snippet_similarity", [_c("a b c d", "a b c e"), _c("", "x"), _c("same", "same")]),
    ("crawler.freshness", "classify_freshness", [_c(1_000_000.0 - 3600, now=1_000_000.0), _c(1_000_000.0 - 86400 * 3, now=1_000_000.0), _c(0.0, now=1_000_000.0 * 100)]),
    ("crawler.recrawl", "compute_recrawl_interval", [_c(0.0), _c(0.2), _c(0.5), _c(0.9), _c(1.0)]),
    ("crawler.recrawl", "update_change_frequency", [_c(0.5, True), _c(0.5, False), _c(0.0, True, alpha=0.5)]),
    ("summarizer.verify", "extract_key_facts", [_c(T1)]),
    ("summarizer.verify", "detect_contradiction", [_c("Revenue was 5 million in 2020.", "Revenue was 9 million in 2020."), _c(T1, "Tensor memory is 256 KB per SM.")]),
    ("summarizer.verify", "self_verify", [_c(T1, "Blackwell keeps accumulators in tensor memory, which is 256 KB per SM. TMA moves tiles and supports multicast."), _c(T1, "Cats are nice.")]),
    ("summarizer.verify", "compute_similarity", [_c("a b c", "a b d"), _c("", "")]),
    ("summarizer.verify", "verify_summary", [_c("https://e.org", "h", T1, "Tensor memory is 256 KB per SM and holds accumulators.", peer_summaries=["Tensor memory holds MMA accumulators, 256 KB per SM."])]),
    ("search.extended", "translate_query_keywords", [_c("검색 엔진", "ko"), _c("hello", "en"), _c("recherche rapide", "fr")]),
    ("crawler.intelligence", "extract_image_alt_texts", [_c('<img src="a.png" alt="A chart of bandwidth"><img alt=""><img src=b alt="x">')]),
    ("credits.timezone_verify", "get_timezone_offset", [_c("Asia/Seoul"), _c("UTC"), _c("Nope/Zone"), _c("America/New_York")]),
    ("security_ext", "sign_webhook_payload", [_c({"b": 1, "a": "x"}, "secret")]),
    ("security_ext", "check_role", [_c("web_search", "reader"), _c("crawl_url", "reader"), _c("crawl_url", "admin"), _c("status", None)]),
    ("search.formatter", "format_fetch_result", [_c(title="T", url="https://u", text="body", is_cached=True, crawled_at=1.0), _c(title="T", url="https://u", text="b", is_cached=False, is_paywall=True)]),
    ("search.formatter", "format_fts_results", [_c(lambda p: _query_result(p)), _c(lambda p: _query_result(p, empty=True))]),
    ("search.formatter", "format_fts_results_json", [_c(lambda p: _query_result(p))]),
    ("data_quality", "cross_reference_results", [_c("tensor memory is 256 KB per SM", lambda p: _ranked(p))]),
    ("dx", "generate_tool_guide", [_c(), _c(format="markdown")]),
    ("search.merge", "merge_results", [_c(lambda p: _fts(p), lambda p: _vec(p), limit=5)]),
]


def _ranked(pkg):
    import importlib

    RR = importlib.import_module(pkg + ".index.ranking").RankedResult
    rows = [("https://a.org/1", "Tensor memory guide", "Tensor memory is 256 KB per SM and holds accumulators", 0.9),
            ("https://a.org/2", "Tensor memory guide (mirror)", "Tensor memory is 256 KB per SM and holds accumulators", 0.8),
            ("https://b.org/x", "TMA multicast", "TMA supports multicast across a thread block cluster", 0.7),
            ("https://b.org/y", "TMA descriptors", "A CUtensorMap descriptor is needed per tensor for TMA", 0.6),
            ("https://c.io/z", "Unrelated", "Cooking pasta takes ten minutes", 0.2)]
    return [RR(doc_id=i, url=u, title=t, snippet=s, bm25_score=sc, freshness_score=0.5, trust_score=0.5, authority_score=0.1, combined_score=sc,
               crawled_at=1_700_000_000.0) for i, (u, t, s, sc) in enumerate(rows)]


def _query_result(pkg, empty=False):
    import importlib

    QR = importlib.import_module(pkg + ".search.query").QueryResult
    import inspect

    rows = [] if empty else _ranked(pkg)
    params = inspect.signature(QR).parameters
    kw = {"results": rows, "total": len(rows), "elapsed_ms": 12.34, "query": "tensor memory"}
    if "source" in params:
        kw["source"] = "local_fts"
    return QR(**{k: v for k, v in kw.items() if k in params})


def _fts(pkg):
    import importlib

    SR = importlib.import_module(pkg + ".index.local_store").SearchResult
    import inspect

    p = inspect.signature(SR).parameters
    base = [dict(doc_id=1, url="https://a.org/1", title="A", snippet="sa", score=3.0, language="en", crawled_at=1.0),
            dict(doc_id=2, url="https://b.org/2", title="B", snippet="sb", score=2.0, language="en", crawled_at=1.0)]
    return [SR(**{k: v for k, v in r.items() if k in p}) for r in base]


def _vec(pkg):
    import importlib

    VR = importlib.import_module(pkg + ".index.vector_store").VectorSearchResult
    return [VR(doc_id="2", url="https://b.org/2", title="B", text_preview="pb", score=0.9), VR(doc_id="3", url="https://c.org/3", title="C", text_preview="pc", score=0.5)]


# ----------------------------------------------------------------------------- round 2: wider input sweeps over the same functions
_URLS = ["https://Example.COM/a/../b/./c?x=1&utm_medium=m&y=2#top", "http://example.com:8080/path;params?q=1", "https://example.com/%7Euser/?a=1&a=2&b=",
         "HTTP://EXAMPLE.COM", "https://example.com/a?", "https://example.com/a#", "https://example.com//double//slash/", "https://example.com/trailing/",
         "https://example.com/index.html?ref=home&id=5", "https://xn--bcher-kva.example/ü?q=ä", "https://example.com/a b c", "https://example.com/?b=2&a=1&c=3",
         "https://user@example.com/", "https://example.com:443", "http://example.com:80", "https://example.com/path?mc_cid=1&mc_eid=2&real=1",
         "https://example.com/?utm_source=a", "ftp://example.com/file", "not a url", "", "https://example.com/%E2%9C%93", "https://EXAMPLE.com/CaseSensitive/Path"]
_HOSTS = ["http://192.168.1.1/", "http://172.16.0.1/", "http://172.32.0.1/", "http://[fe80::1]/", "http://[fd00::1]/", "http://0.0.0.0/", "http://2130706433/",
          "http://example.com@127.0.0.1/", "https://example.org:8443/x", "http://metadata.google.internal/", "https://sub.example.co.uk/a?b=c", "javascript:alert(1)",
          "file:///etc/passwd", "http://100.64.0.1/", "http://198.18.0.1/", "https://1.1.1.1/"]
_QUERIES = ["python asyncio tutorial", "how do i install cuda on ubuntu", "best GPU 2024", "\"exact phrase\" -excluded", "site:github.com tensor cores", "C++ templates",
            "  spaces   everywhere  ", "UPPER lower MiXeD", "日本語のクエリ", "한국어 질문 입니다", "emoji 🚀 query", "a", "the of and", "x" * 300, "what is the capital of france?",
            "login facebook", "buy cheap laptop online", "weather tomorrow", "define: tensor", "error: undefined reference to `main'"]
_TEXTS = [T1, T2, T3, T4, "", "short", "One sentence only.", "Numbers 123 456.789 and symbols #!$%", "Line one\nLine two\n\nParagraph two starts here. It has two sentences.",
          "ALL CAPS TEXT WITH MANY WORDS TO SEE HOW TOKENISATION BEHAVES", "mixed 한글 and English words together in one sentence", "repeat " * 50,
          "Bonjour le monde, ceci est une phrase en français avec des accents éàç.", "Hola mundo, esta es una oración en español con eñe.",
          "Привет мир, это предложение на русском языке.", "这是一个中文句子，用来测试分词。", "مرحبا بالعالم هذه جملة عربية"]

CASES += [
    ("crawler.dedup", "normalize_url", [_c(u) for u in _URLS]),
    ("security", "validate_url", [_c(u) for u in _HOSTS]),
    ("search.query", "_sanitize_fts_query", [_c(q) for q in _QUERIES]),
    ("search.passage", "classify_intent", [_c(q) for q in _QUERIES]),
    ("search.nlp", "expand_query", [_c(q) for q in _QUERIES[:12]]),
    ("search.nlp", "parse_natural_query", [_c(q) for q in _QUERIES[:12]]),
    ("search.cjk", "tokenize_query_cjk", [_c(q) for q in _QUERIES[:12]]),
    ("search.cjk", "is_cjk_text", [_c(t) for t in _TEXTS]),
    ("search.quality", "extract_temporal_hint", [_c(q) for q in ("latest news", "this week in ai", "yesterday's game", "2019 budget", "last year sales", "recent papers", "old maps")]),
    ("crawler.simhash", "simhash", [_c(t) for t in _TEXTS]),
    ("search.passage", "split_passages", [_c(t) for t in _TEXTS]),
    ("search.passage", "select_best_passage", [_c(t, "tensor memory sentence") for t in _TEXTS[:10]]),
    ("index.distributed", "extract_keywords", [_c(t) for t in _TEXTS]),
    ("search.passage", "title_match_score", [_c(t[:60], "tensor memory") for t in _TEXTS[:8]]),
    ("search.passage", "url_path_score", [_c(u, "example path index") for u in _URLS[:10]]),
    ("p2p.protocol", "keyword_to_dht_key", [_c(w) for w in ("rust", "RUST", " rust ", "c++", "日本", "a" * 200)]),
    ("p2p.protocol", "url_to_dht_key", [_c(u) for u in _URLS[:8]]),
    ("hashing", "content_hash", [_c(t) for t in _TEXTS[:8]]),
    ("crawler.simhash", "hamming_distance", [_c(a, b) for a, b in ((1, 2), (0xFFFF, 0xFFFE), (2 ** 63, 0), (12345678901234567890 % 2 ** 64, 987654321))]),
    ("summarizer.verify", "compute_similarity", [_c(a, b) for a, b in ((T1, T1[:200]), (T1, T4), ("a b c d e", "e d c b a"), ("x", ""))]),
    ("search.cross_validate", "snippet_similarity", [_c(a, b) for a, b in ((T1[:100], T1[50:150]), ("one two three", "three two one"), ("", ""))]),
    ("data_quality", "compute_trust_grade", [_c(x / 20) for x in range(0, 21)]),
    ("crawler.recrawl", "compute_recrawl_interval", [_c(x / 10) for x in range(0, 11)]),
    ("index.ranking", "freshness_score", [_c(1_000_000.0 - d * 86400.0, now=1_000_000.0) for d in (0, 1, 3, 7, 14, 30, 90, 365)]),
    ("index.ranking", "normalize_bm25", [_c(float(s), max_score=float(m)) for s, m in ((1, 1), (5, 10), (10, 5), (0.1, 20), (100, 1))]),
    ("credits.scheduling", "is_off_peak_at", [_c(hour=h) for h in range(0, 24, 3)]),
    ("version_check", "_parse_version", [_c(v) for v in ("0.1.0", "1.2.3rc1", "2024.01.15", "v2", "abc", "1.0.0+build.5", "10.20.30.40")]),
    ("version_check", "is_newer", [_c(a, b) for a, b in (("1.2.4", "1.2.3"), ("1.2.3", "1.2.3"), ("2.0", "1.99.99"), ("0.9", "1.0"))]),
    ("crawler.pdf", "is_pdf_url", [_c(u) for u in ("https://x.org/a.pdf", "https://x.org/a.PDF?dl=1", "https://x.org/pdf", "https://x.org/a.pdf#page=2", "")]),
    ("p2p.peer_profile", "_percentile", [_c(v, q) for v, q in (([1.0, 2.0, 3.0], 50), ([1.0], 99), ([4.0, 1.0, 3.0, 2.0], 25), ([10.0, 20.0], 95))]),
    ("crawler.freshness", "classify_freshness", [_c(1_000_000.0 - a, now=1_000_000.0) for a in (0, 1800, 3600, 7200, 86400, 172800, 604800, 1209600)]),
]
