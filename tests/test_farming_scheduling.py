"""credits/farming.py, credits/scheduling.py, credits/timezone_verify.py — deterministic ``now=`` overrides, in-memory SQLite."""
import pytest

from infomesh_b200.credits import farming as F
from infomesh_b200.credits import scheduling as S
from infomesh_b200.credits import timezone_verify as TZ

T0 = 1_000_000.0


def test_probation_window_and_multiplier():
    d = F.FarmingDetector()
    assert d.probation_remaining("ghost") == F.PROBATION_HOURS          # unknown node: full probation
    d.register_node("p", now=T0)
    d.register_node("p", now=T0 + 999)                                   # re-registration does not reset the clock
    assert d.probation_remaining("p", now=T0 + 6 * 3600) == pytest.approx(18.0)
    assert d.is_on_probation("p", now=T0 + 23 * 3600) and not d.is_on_probation("p", now=T0 + 25 * 3600)
    assert d.credit_multiplier("p", now=T0 + 3600) == 0.5 and d.credit_multiplier("p", now=T0 + 25 * 3600) == 1.0


def test_hourly_rate_limits_per_action():
    d = F.FarmingDetector()
    for i in range(F.MAX_LLM_PER_HOUR):
        d.log_action("p", "llm_own", now=T0 + i * 50.0 + (i % 7))        # irregular spacing
    assert d.actions_in_last_hour("p", "llm_own", now=T0 + 3500) == F.MAX_LLM_PER_HOUR
    assert d.is_rate_limited("p", "llm_own", now=T0 + 3500) and not d.is_rate_limited("p", "crawl", now=T0 + 3500)
    assert not d.is_rate_limited("p", "llm_own", now=T0 + 3 * 3600)      # window slides


def test_regular_interval_detection_needs_ten_samples_and_low_variance():
    d = F.FarmingDetector()
    for i in range(9):
        d.log_action("bot", "crawl", now=T0 + i * 30.0)
    assert not d.detect_regular_intervals("bot", "crawl", now=T0 + 300)
    d.log_action("bot", "crawl", now=T0 + 270.0)
    assert d.detect_regular_intervals("bot", "crawl", now=T0 + 300)
    gaps = [3, 40, 7, 90, 15, 60, 2, 33, 71, 11, 5]
    t = T0
    for g in gaps:
        t += g
        d.log_action("human", "crawl", now=t)
    assert not d.detect_regular_intervals("human", "crawl", now=t + 1)


def test_burst_detection_and_blocking_after_three_anomalies():
    d = F.FarmingDetector()
    d.register_node("p", now=T0 - 48 * 3600)
    for i in range(F.BURST_THRESHOLD):
        d.log_action("p", "crawl", now=T0 + i * (1.0 + (i % 5) * 0.7))
    assert d.detect_burst("p", "crawl", now=T0 + 120)
    c1 = d.check("p", "crawl", now=T0 + 120)
    assert c1.verdict == F.FarmingVerdict.SUSPICIOUS and "burst" in c1.detail
    d.record_anomaly("p", "manual", now=T0 + 121)
    d.record_anomaly("p", "manual", now=T0 + 122)
    assert d.is_blocked("p") and d.credit_multiplier("p") == 0.0
    assert d.check("p", "crawl", now=T0 + 130).verdict == F.FarmingVerdict.BLOCKED
    assert len(d.get_anomaly_history("p")) >= 3
    d.unblock("p")
    assert not d.is_blocked("p") and d.check("p", "search", now=T0 + 10 * 3600).verdict == F.FarmingVerdict.CLEAN


def test_check_verdict_precedence_probation_vs_clean():
    d = F.FarmingDetector()
    assert d.check("new", "crawl", now=T0).verdict == F.FarmingVerdict.PROBATION
    assert d.check("new", "crawl", now=T0 + 30 * 3600).verdict == F.FarmingVerdict.CLEAN


def test_prune_old_actions():
    d = F.FarmingDetector()
    d.log_action("p", "crawl", now=1.0)
    d.log_action("p", "crawl")
    assert d.prune_old_actions() == 1


# ------------------------------------------------------------------ scheduling
def node(pid, tz="UTC", llm=True, trust=0.5, start=23, end=7, ip=""):
    return S.NodeScheduleInfo(pid, start, end, tz, llm, trust, ip)


def test_off_peak_window_wraps_midnight():
    assert S.is_off_peak_at(hour=23) and S.is_off_peak_at(hour=3) and not S.is_off_peak_at(hour=7)
    assert S.is_off_peak_at(hour=13, start=12, end=14) and not S.is_off_peak_at(hour=14, start=12, end=14)
    assert S.parse_hhmm("23:00", 1) == 23 and S.parse_hhmm("7", 1) == 7 and S.parse_hhmm("xx", 5) == 5 and S.parse_hhmm("99:00", 5) == 5


def test_schedule_prefers_off_peak_then_trust():
    sch = S.EnergyAwareScheduler()
    nodes = [node("day", start=1, end=2, trust=0.9), node("night-lo", trust=0.3), node("night-hi", trust=0.8), node("no-llm", llm=False, trust=1.0)]
    d = sch.schedule_llm_task(nodes, now_override_hour=3)
    assert d.target_peer_id == "night-hi" and d.is_off_peak and d.credit_multiplier == S.OFF_PEAK_MULTIPLIER
    d = sch.schedule_llm_task(nodes, now_override_hour=12)
    assert d.target_peer_id == "day" and not d.is_off_peak and d.credit_multiplier == 1.0
    assert sch.schedule_llm_task([node("x", llm=False)], now_override_hour=3) is None


def test_schedule_batch_round_robin():
    sch = S.EnergyAwareScheduler()
    nodes = [node("a", trust=0.9), node("b", trust=0.5), node("peak", start=1, end=2)]
    out = sch.schedule_batch(nodes, 5, now_override_hour=2)
    assert [d.target_peer_id for d in out] == ["a", "b", "a", "b", "a"] and all(d.is_off_peak for d in out)
    assert sch.schedule_batch(nodes, 0, now_override_hour=2) == [] and sch.schedule_batch([], 3) == []


def test_off_peak_claim_rejected_when_timezone_contradicts_ip():
    sch = S.EnergyAwareScheduler()
    liar = node("liar", tz="America/New_York", ip="211.5.5.5")        # Korean address space, claims US east coast
    honest = node("honest", tz="Asia/Seoul", ip="211.6.6.6", trust=0.1)
    d = sch.schedule_llm_task([liar, honest], now_override_hour=3)
    assert d.target_peer_id == "honest"


# ------------------------------------------------------------------ timezone verification
def test_timezone_offsets_and_ip_estimates():
    assert TZ.get_timezone_offset("Asia/Seoul") == 9.0 and TZ.get_timezone_offset("Not/AZone") == 0.0
    assert TZ.estimate_offset_from_ip("211.1.1.1") == 9.0 and TZ.estimate_offset_from_ip("garbage") is None
    assert TZ.estimate_offset_from_ip("250.0.0.1") is None


def test_verify_timezone_wraps_around_the_date_line():
    c = TZ.verify_timezone("p", "Pacific/Auckland", "150.1.1.1")       # +12/+13 vs +10
    assert c.offset_diff_hours <= 3
    far = TZ.verify_timezone("p", "Pacific/Honolulu", "150.1.1.1")    # -10 vs +10 -> 4 h across the date line
    assert far.offset_diff_hours == 4.0 and not far.plausible


def test_consistency_tracker_flags_frequent_changes_and_forgets_old_claims():
    t = TZ.TimezoneConsistencyTracker()
    zones = ["Asia/Seoul", "Europe/Paris", "Asia/Seoul", "America/New_York"]
    recs = [t.record_claim("p", z, now=T0 + i * 60) for i, z in enumerate(zones)]
    assert [r.changes_in_24h for r in recs] == [0, 1, 2, 3] and recs[-1].suspicious and recs[-1].unique_timezones == 3
    later = t.record_claim("p", "Asia/Seoul", now=T0 + 72 * 3600)
    assert later.claim_count == 1 and not later.suspicious
