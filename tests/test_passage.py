"""search/passage.py — passage splitting, coverage+density scoring, best-passage selection, highlighting, intent."""
from infomesh_b200.search import passage as P


def test_tokenize_lowercases_and_keeps_unicode_words():
    # Latin / digits, Hangul and CJK ideographs form tokens (kana is not in the reference's character classes either)
    assert P.tokenize("Hello, World! 한국어 漢字 テスト") == ["hello", "world", "한국어", "漢字"]


def test_split_passages_respects_paragraphs_and_max_length():
    text = ("First paragraph about python asyncio which is long enough to stand alone here.\n\n"
            "Second paragraph about rust and tokio that is also long enough to be kept alone.")
    ps = P.split_passages(text, max_length=200)
    assert len(ps) == 2 and ps[0].startswith("First") and ps[1].startswith("Second")
    assert P.split_passages("") == [] and P.split_passages("   \n ") == []


def test_split_passages_folds_short_paragraphs_into_predecessor():
    text = "A long enough opening paragraph that easily passes the forty character minimum.\n\nShort tail."
    ps = P.split_passages(text)
    assert len(ps) == 1 and ps[0].endswith("Short tail.")


def test_split_passages_breaks_long_paragraph_on_sentences():
    sent = "This sentence talks about search engines and ranking. "
    ps = P.split_passages(sent * 30, max_length=200)
    assert len(ps) > 3 and all(len(p) <= 260 for p in ps)


def test_score_passage_is_coverage_plus_tenth_density():
    s = P.score_passage("python asyncio python", ["python", "rust"])
    assert abs(s - (0.5 + 0.1 * (2 / 3))) < 1e-9
    assert P.score_passage("", ["x"]) == 0.0 and P.score_passage("text", []) == 0.0


def test_rank_passages_best_first_with_offsets():
    text = ("Cooking pasta requires boiling water and a pinch of salt for the best results.\n\n"
            "Python asyncio provides an event loop, coroutines and tasks for concurrent programs.")
    ranked = P.rank_passages(text, "python asyncio")
    assert ranked[0].text.startswith("Python") and ranked[0].score > ranked[1].score
    assert text[ranked[0].start:ranked[0].start + 6] == "Python"


def test_select_best_passage_and_fallbacks():
    text = ("Cooking pasta requires boiling water and a pinch of salt for the best results.\n\n"
            "Python asyncio provides an event loop, coroutines and tasks for concurrent programs.")
    assert P.select_best_passage(text, "asyncio event loop").startswith("Python")
    assert P.select_best_passage(text, "") == text[:200]
    assert P.select_best_passage(text, "zebra") == text[:200]
    assert P.select_best_passage("", "q") == ""


def test_highlight_terms_wraps_whole_words_case_insensitively():
    assert P.highlight_terms("Python and pythonic", ["python"]) == "<b>Python</b> and pythonic"
    assert P.highlight_terms("text", []) == "text"


def test_title_and_url_scores():
    assert P.title_match_score("Python Asyncio Guide", ["python", "rust"]) == 0.5
    assert P.title_match_score("", ["x"]) == 0.0
    assert P.url_path_score("https://ex.org/docs/python-asyncio.html", ["asyncio", "rust"]) == 0.5
    assert P.url_path_score("https://ex.org/", ["x"]) == 0.0


def test_classify_intent():
    assert P.classify_intent("github login") == P.QueryIntent.NAVIGATIONAL
    assert P.classify_intent("example.com") == P.QueryIntent.NAVIGATIONAL
    assert P.classify_intent("download python installer") == P.QueryIntent.TRANSACTIONAL
    assert P.classify_intent("how does bm25 work") == P.QueryIntent.INFORMATIONAL
    assert P.classify_intent("") == P.QueryIntent.INFORMATIONAL
