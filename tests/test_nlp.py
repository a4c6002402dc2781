"""search/nlp.py — stop words, synonym expansion, typo correction, natural-language filters, related searches
(reference tests/test_nlp.py shape: pure unit, no I/O)."""
import time

from infomesh_b200.search import nlp


def test_stop_words_default_to_english_and_are_case_insensitive():
    assert "the" in nlp.get_stop_words() and nlp.get_stop_words("xx") == nlp.get_stop_words("en")
    assert nlp.remove_stop_words(["The", "quick", "and", "fox"]) == ["quick", "fox"]


def test_stop_words_other_language_tables_exist():
    for lang in ("ko", "ja", "de", "fr", "es"):
        assert isinstance(nlp.get_stop_words(lang), frozenset)


def test_expand_query_adds_synonyms_without_repeating_query_words():
    extra = nlp.expand_query("database error")
    assert "db" in extra and "exception" in extra
    assert "database" not in extra and "error" not in extra
    assert len(extra) <= 6


def test_expand_query_respects_max_expansions():
    assert len(nlp.expand_query("error api database server config", max_expansions=1)) <= 2
    assert nlp.expand_query("zzzz qqqq") == []


def test_edit_distance_basics():
    assert nlp.edit_distance("kitten", "sitting") == 3
    assert nlp.edit_distance("", "abc") == 3 and nlp.edit_distance("abc", "abc") == 0
    assert nlp.edit_distance("ab", "ba") == 2


def test_did_you_mean_fixes_one_typo_and_keeps_known_words():
    vocab = ["python", "asyncio", "tutorial", "search"]
    assert nlp.did_you_mean("pyhton tutorial", vocab) == ["python tutorial"]
    assert nlp.did_you_mean("python tutorial", vocab) == []


def test_did_you_mean_ignores_far_words_and_limits_suggestions():
    vocab = ["alpha", "alphb", "alphc", "alphd"]
    assert nlp.did_you_mean("zzzzzzzz", vocab) == []
    assert len(nlp.did_you_mean("alphx betx gamx", ["alpha", "beta", "gamma"], max_suggestions=2)) == 2


def test_parse_natural_query_dates():
    now = time.time()
    p = nlp.parse_natural_query("python news last 3 days")
    assert p.cleaned_query == "python news" and abs(p.date_from - (now - 3 * 86400)) < 5
    p = nlp.parse_natural_query("recent rust releases")
    assert p.cleaned_query == "rust releases" and abs(p.date_from - (now - 7 * 86400)) < 5
    assert nlp.parse_natural_query("plain query").date_from is None


def test_parse_natural_query_site_and_language():
    p = nlp.parse_natural_query("asyncio guide site:docs.python.org in korean")
    assert p.include_domains == ["docs.python.org"] and p.language == "ko"
    assert p.cleaned_query == "asyncio guide" and p.original_query.endswith("in korean")
    p = nlp.parse_natural_query("tutorials from example.com")
    assert p.include_domains == ["example.com"] and p.cleaned_query == "tutorials"


def test_related_search_tracker_counts_cooccurrence():
    t = nlp.RelatedSearchTracker()
    for q in ("python asyncio", "python asyncio tutorial", "python typing", "rust tokio"):
        t.record(q)
    rel = t.related("python")
    assert rel[0] == "asyncio" and "tokio" not in rel
    assert t.related("unknown") == []


def test_related_search_tracker_prunes_when_over_budget():
    t = nlp.RelatedSearchTracker(max_pairs=10)
    for i in range(30):
        t.record(f"a{i} b{i} c{i}")
    assert len(t._pairs) <= 10
