"""credits/ledger.py (weights, caps, tiers, grace / debt), credits/github_identity.py, index/ranking.py."""
import time
from types import SimpleNamespace
from unittest.mock import patch

import pytest

from infomesh_b200.credits import github_identity as GH
from infomesh_b200.credits.ledger import CreditLedger, is_off_peak
from infomesh_b200.credits.types import (ACTION_WEIGHTS, DEBT_COST_MULTIPLIER, GRACE_PERIOD_HOURS, ActionType,
                                          ContributionTier, CreditState)
from infomesh_b200.index import ranking as R
from infomesh_b200.p2p.keys import KeyPair


# ------------------------------------------------------------------ ledger
def test_action_weights_and_off_peak_only_boosts_llm():
    led = CreditLedger()
    assert led.record_action(ActionType.CRAWL, 3) == 3.0
    assert led.record_action(ActionType.CRAWL, 1, off_peak=True) == ACTION_WEIGHTS[ActionType.CRAWL]
    assert led.record_action(ActionType.LLM_SUMMARIZE_PEER, 2, off_peak=True) == 2.0 * 2 * 1.5
    with pytest.raises(ValueError):
        led.record_action(ActionType.CRAWL, 0)
    assert led.total_earned() == 3.0 + 1.0 + 6.0 and dict(led.earnings_by_action())["llm_peer"] == 6.0


def test_llm_share_is_capped_at_sixty_percent_of_the_score():
    led = CreditLedger()
    led.record_action(ActionType.CRAWL, 40)                       # 40 non-LLM
    led.record_action(ActionType.LLM_SUMMARIZE_OWN, 1000)         # 1500 LLM
    assert led.contribution_score() == pytest.approx(40 + 40 * 0.6 / 0.4)
    led2 = CreditLedger()
    led2.record_action(ActionType.CRAWL, 100)
    led2.record_action(ActionType.LLM_SUMMARIZE_OWN, 10)          # 15 / 115 < 60 %: untouched
    assert led2.contribution_score() == pytest.approx(115.0)


def test_tiers_and_search_cost():
    led = CreditLedger()
    assert led.tier() == ContributionTier.TIER_1 and led.search_cost() == 0.100
    led.record_action(ActionType.CRAWL, 100)
    assert led.tier() == ContributionTier.TIER_2 and led.search_cost() == 0.050
    led.record_action(ActionType.GIT_DOCS, 1)
    assert led.tier() == ContributionTier.TIER_3 and led.search_cost() == 0.033


def test_spend_never_blocks_grace_then_debt_then_recovery():
    led = CreditLedger()
    led.record_action(ActionType.CRAWL, 1)
    with pytest.raises(ValueError):
        led.spend(0)
    assert led.spend(0.4) and led.credit_state() == CreditState.NORMAL and led.balance() == pytest.approx(0.6)
    assert led.spend(5.0)                                         # overdraw: allowed, starts the grace clock
    now = time.time()
    assert led.credit_state(now=now) == CreditState.GRACE and led.debt_amount() == pytest.approx(4.4)
    assert 71.9 < led.grace_remaining_hours(now=now) <= GRACE_PERIOD_HOURS
    late = now + (GRACE_PERIOD_HOURS + 1) * 3600
    assert led.credit_state(now=late) == CreditState.DEBT and led.grace_remaining_hours(now=late) is None
    allow = led.search_allowance(now=late)
    assert allow.search_cost == pytest.approx(0.100 * DEBT_COST_MULTIPLIER) and allow.debt_amount == pytest.approx(4.4)
    led.record_action(ActionType.CRAWL, 10)
    assert led.credit_state(now=late) == CreditState.NORMAL and led.search_allowance().grace_remaining_hours is None


def test_signed_entries_and_stats(tmp_path):
    kp = KeyPair.generate()
    led = CreditLedger(tmp_path / "c.db")
    led.owner_email = "me@example.com"
    led.record_action(ActionType.CRAWL, 2, note="n1", key_pair=kp)
    led.record_action(ActionType.QUERY_PROCESS, 4)
    signed = led.signed_entries()
    assert len(signed) == 1 and signed[0].note == "n1" and len(signed[0].signature) == 128
    st = led.stats()
    assert st.total_earned == 4.0 and st.total_spent == 0.0 and len(led.recent_entries(limit=1)) == 1
    led.close()
    again = CreditLedger(tmp_path / "c.db")
    assert again.total_earned() == 4.0
    again.close()


def test_is_off_peak_window():
    assert is_off_peak(hour=23) and is_off_peak(hour=6) and not is_off_peak(hour=12)
    assert is_off_peak(hour=13, start=12, end=14)


# ------------------------------------------------------------------ github identity
def test_email_validation_and_resolution_order():
    assert GH.is_valid_email("a.b+c@ex-ample.org") and not GH.is_valid_email("no-at") and not GH.is_valid_email("")
    cfg = SimpleNamespace(node=SimpleNamespace(github_email="cfg@example.com"))
    assert GH.resolve_github_email(cfg) == "cfg@example.com"
    cfg.node.github_email = ""
    with patch.object(GH, "detect_git_email", return_value="git@example.com"):
        assert GH.resolve_github_email(cfg) == "git@example.com"


def test_detect_git_email_handles_missing_git_and_bad_output():
    with patch("subprocess.run", side_effect=FileNotFoundError):
        assert GH.detect_git_email() is None
    with patch("subprocess.run", return_value=SimpleNamespace(returncode=0, stdout="not-an-email\n")):
        assert GH.detect_git_email() is None
    with patch("subprocess.run", return_value=SimpleNamespace(returncode=0, stdout="dev@example.com\n")):
        assert GH.detect_git_email() == "dev@example.com"


def test_first_start_checks_messages():
    cfg = SimpleNamespace(node=SimpleNamespace(github_email=""))
    out = []
    with patch.object(GH, "detect_git_email", return_value=None):
        assert GH.run_first_start_checks(cfg, interactive=False, echo=out.append) is None
        assert "not connected" in out[0]
        assert GH.run_first_start_checks(cfg, interactive=True, prompt=lambda _: "", echo=out.append) is None
    assert "linked to this account" in GH.format_startup_message("x@y.org")


# ------------------------------------------------------------------ ranking
def cand(url, bm25, age_days=0.0, trust=0.5, authority=0.0, title=0.0, path=0.0, now=1_000_000.0):
    return R.RawCandidate(doc_id=url, url=url, title="T", snippet="s", bm25_raw=bm25, crawled_at=now - age_days * 86400,
                          trust=trust, authority=authority, title_match=title, url_path=path, peer_id=None)


def test_freshness_decay_half_life_floor_and_future():
    now = 1_000_000.0
    assert R.freshness_score(now, now=now) == 1.0 and R.freshness_score(now + 999, now=now) == 1.0
    assert R.freshness_score(now - R.FRESHNESS_HALF_LIFE_SECONDS, now=now) == pytest.approx(0.5)
    assert R.freshness_score(0.0, now=1e12) == R.MIN_FRESHNESS


def test_bm25_saturation_normalisation():
    assert R.normalize_bm25(5.0, max_score=5.0) == 0.5 and R.normalize_bm25(0, max_score=5) == 0.0 and R.normalize_bm25(3, max_score=0) == 1.0     # k = 0 saturates immediately (reference ranking.py:99-101)
    assert R.normalize_bm25(10, max_score=5) > R.normalize_bm25(5, max_score=5)


def test_combined_score_weights_sum_to_one():
    w = R.WEIGHT_BM25 + R.WEIGHT_FRESHNESS + R.WEIGHT_TRUST + R.WEIGHT_AUTHORITY + R.WEIGHT_TITLE_MATCH + R.WEIGHT_URL_PATH
    assert w == pytest.approx(1.0) and R.combined_score(1, 1, 1, 1, title_match=1, url_path=1) == pytest.approx(1.0)


def test_rank_results_orders_by_combined_signal_and_honours_weights():
    now = 1_000_000.0
    cs = [cand("old-strong", 10.0, age_days=400), cand("fresh-weak", 2.0, age_days=0), cand("titled", 6.0, age_days=1, title=1.0, path=1.0)]
    out = R.rank_results(cs, now=now)
    assert out[0].url == "titled" and {r.url for r in out} == {"old-strong", "fresh-weak", "titled"}
    assert out[0].bm25_score == round(6 / 16, 6) and 0 < out[0].freshness_score < 1
    only_bm25 = R.rank_results(cs, now=now, weights=dict(w_bm25=1.0, w_fresh=0, w_trust=0, w_authority=0, w_title=0, w_url=0))
    assert [r.url for r in only_bm25] == ["old-strong", "titled", "fresh-weak"]
    assert R.rank_results([], now=now) == [] and len(R.rank_results(cs, limit=2, now=now)) == 2


def test_rank_local_results_uses_query_tokens_and_survives_bad_authority():
    now = time.time()
    rows = [SimpleNamespace(doc_id=1, url="https://ex.org/python-asyncio", title="Python asyncio", snippet="s", score=3.0, crawled_at=now),
            SimpleNamespace(doc_id=2, url="https://ex.org/other", title="Other", snippet="s", score=3.0, crawled_at=now)]

    def bad_authority(url):
        raise RuntimeError("db locked")

    out = R.rank_local_results(rows, query_tokens=["python", "asyncio"], authority_fn=bad_authority)
    assert [r.doc_id for r in out] == [1, 2] and out[0].title_match_score == 1.0 and out[0].url_path_score == 1.0
    assert out[0].authority_score == 0.0
