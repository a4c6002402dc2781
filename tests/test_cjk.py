"""search/cjk.py — script detection, n-gram tokenisation, tokenizer recommendation, Korean segmentation."""
from infomesh_b200.search import cjk


def test_is_cjk_text_threshold():
    assert cjk.is_cjk_text("한국어 검색 엔진")
    assert cjk.is_cjk_text("漢字テスト")
    assert not cjk.is_cjk_text("plain english sentence")
    assert not cjk.is_cjk_text("") and not cjk.is_cjk_text("   ")
    assert not cjk.is_cjk_text("mostly english with 한 char", threshold=0.3)


def test_bigrams_and_trigrams_keep_latin_runs_whole():
    assert cjk.cjk_bigrams("검색엔진 v2") == ["검색", "색엔", "엔진", "v2"]
    assert cjk.cjk_trigrams("検索エンジン") == ["検索エ", "索エン", "エンジ", "ンジン"]
    assert cjk.cjk_bigrams("日") == ["日"]          # run shorter than n stays intact


def test_recommend_tokenizer():
    assert cjk.recommend_tokenizer("分散型検索エンジンの設計") == "trigram"
    assert cjk.recommend_tokenizer("distributed search engine design") == "unicode61"


def test_tokenize_query_rewrites_only_cjk_queries():
    assert cjk.tokenize_query_cjk("python asyncio") == "python asyncio"
    assert cjk.tokenize_query_cjk("검색엔진") == "검색 색엔 엔진"


def test_segment_korean_short_words_whole_long_words_bigrams():
    assert cjk.segment_korean("검색 엔진 python3") == ["검색", "엔진", "python3"]
    assert cjk.segment_korean("분산검색엔진") == ["분산", "산검", "검색", "색엔", "엔진"]


def test_segment_chinese_falls_back_to_bigrams_without_jieba():
    toks = cjk.segment_chinese("分布式搜索")
    assert "".join(t for t in toks if len(t) == 1) or all(len(t) >= 1 for t in toks)
    assert "".join(dict.fromkeys("".join(toks))) .startswith("分")


def test_detect_script():
    assert cjk.detect_script("한국어") == "hangul" and cjk.detect_script("漢字") == "cjk"
    assert cjk.detect_script("カタカナ") == "kana" and cjk.detect_script("ภาษาไทย") == "thai"
    assert cjk.detect_script("العربية") == "arabic" and cjk.detect_script("हिन्दी") == "devanagari"
    assert cjk.detect_script("hello") == "latin"
