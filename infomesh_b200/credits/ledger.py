"""Local contribution ledger (SQLite): earn = weight x quantity x multiplier, spend never blocks, a 72 h grace
window then x2 search cost in debt, LLM credits capped at 60 % of the contribution score, tiered search cost.
Tables, hashing/signing canon and formulas per reference infomesh/credits/ledger.py:81-652."""
from __future__ import annotations

import contextlib
import datetime
import time
from pathlib import Path

from infomesh_b200.credits.scheduling import (DEFAULT_OFF_PEAK_END, DEFAULT_OFF_PEAK_START, OFF_PEAK_MULTIPLIER,
                                               is_off_peak_at)
from infomesh_b200.credits.types import (_LLM_ACTIONS, ACTION_WEIGHTS, DEBT_COST_MULTIPLIER,
                                          GIT_CONTRIBUTION_LEGACY, GRACE_PERIOD_HOURS, LLM_CREDIT_CAP_RATIO,
                                          TIER_THRESHOLDS, ActionType, ContributionTier, CreditEntry, CreditState,
                                          LedgerStats, SearchAllowance)
from infomesh_b200.db import SQLiteStore
from infomesh_b200.hashing import content_hash
from infomesh_b200.types import KeyPairLike
from infomesh_b200.utils.log import get_logger

logger = get_logger(__name__)

__all__ = ["ACTION_WEIGHTS", "ActionType", "ContributionTier", "CreditEntry", "CreditLedger", "CreditState",
           "DEBT_COST_MULTIPLIER", "GIT_CONTRIBUTION_LEGACY", "GRACE_PERIOD_HOURS", "LLM_CREDIT_CAP_RATIO",
           "LedgerStats", "SearchAllowance", "TIER_THRESHOLDS", "is_off_peak"]

_LLM_VALUES = tuple(a.value for a in _LLM_ACTIONS)
_ENTRY_COLS = "entry_id, action, quantity, weight, multiplier, credits, timestamp, note, entry_hash, signature"


def _entry_canonical(action: str, quantity: float, weight: float, multiplier: float, credits: float,
                     timestamp: float, note: str) -> bytes:
    """``action|quantity|weight|multiplier|credits|timestamp|note`` — the bytes that are hashed and signed."""
    return f"{action}|{quantity}|{weight}|{multiplier}|{credits}|{timestamp}|{note}".encode()


def _score_to_tier(score: float) -> ContributionTier:
    for floor, tier, _ in TIER_THRESHOLDS:
        if score >= floor:
            return tier
    return ContributionTier.TIER_1


def _tier_search_cost(tier: ContributionTier) -> float:
    return next((cost for _, t, cost in TIER_THRESHOLDS if t == tier), 0.100)


def is_off_peak(*, hour: int | None = None, start: int | None = None, end: int | None = None) -> bool:
    if hour is None:
        hour = datetime.datetime.now().hour  # noqa: DTZ005 — local wall clock is the point
    return is_off_peak_at(hour=hour, start=DEFAULT_OFF_PEAK_START if start is None else start,
                          end=DEFAULT_OFF_PEAK_END if end is None else end)


class CreditLedger(SQLiteStore):
    _SCHEMA = """
        CREATE TABLE IF NOT EXISTS credit_entries (
            entry_id INTEGER PRIMARY KEY AUTOINCREMENT, action TEXT NOT NULL, quantity REAL NOT NULL,
            weight REAL NOT NULL, multiplier REAL NOT NULL DEFAULT 1.0, credits REAL NOT NULL,
            timestamp REAL NOT NULL, note TEXT NOT NULL DEFAULT '', entry_hash TEXT NOT NULL DEFAULT '',
            signature TEXT NOT NULL DEFAULT '', owner_email TEXT NOT NULL DEFAULT '');
        CREATE TABLE IF NOT EXISTS credit_spending (
            spend_id INTEGER PRIMARY KEY AUTOINCREMENT, amount REAL NOT NULL,
            reason TEXT NOT NULL DEFAULT 'search', timestamp REAL NOT NULL);
        CREATE TABLE IF NOT EXISTS credit_grace (id INTEGER PRIMARY KEY CHECK (id = 1), grace_start REAL);
        INSERT OR IGNORE INTO credit_grace (id, grace_start) VALUES (1, NULL);
        CREATE INDEX IF NOT EXISTS idx_entries_action ON credit_entries(action);
        CREATE INDEX IF NOT EXISTS idx_entries_ts ON credit_entries(timestamp);
    """

    def __init__(self, db_path: Path | str | None = None, *, owner_email: str = ""):
        self._owner_email = owner_email
        super().__init__(db_path)

    def _post_init(self) -> None:
        have = {r[1] for r in self._conn.execute("PRAGMA table_info(credit_entries)")}
        for col in ("entry_hash", "signature", "owner_email"):      # ledgers written before signing existed
            if col not in have:
                self._conn.execute(f"ALTER TABLE credit_entries ADD COLUMN {col} TEXT NOT NULL DEFAULT ''")
        self._conn.commit()

    @property
    def owner_email(self) -> str:
        return self._owner_email

    @owner_email.setter
    def owner_email(self, value: str) -> None:
        self._owner_email = value

    # ------------------------------------------------------------------ earn / spend
    def record_action(self, action: ActionType, quantity: float = 1.0, *, off_peak: bool = False, note: str = "",
                      key_pair: KeyPairLike | None = None) -> float:
        if quantity <= 0:
            raise ValueError(f"quantity must be positive, got {quantity}")
        quantity = float(quantity)      # SQLite hands back REAL: keep the signed canon identical after a round trip
        weight = ACTION_WEIGHTS[action]
        mult = OFF_PEAK_MULTIPLIER if (off_peak and action in _LLM_ACTIONS) else 1.0
        earned = weight * quantity * mult
        now = time.time()
        canon = _entry_canonical(action.value, quantity, weight, mult, earned, now, note)
        sig = key_pair.sign(canon).hex() if key_pair is not None else ""
        with self._lock:
            self._conn.execute(
                "INSERT INTO credit_entries (action, quantity, weight, multiplier, credits, timestamp, note, "
                "entry_hash, signature, owner_email) VALUES (?, ?, ?, ?, ?, ?, ?, ?, ?, ?)",
                (action.value, quantity, weight, mult, earned, now, note, content_hash(canon), sig,
                 self._owner_email))
            self._conn.commit()
        if self.balance() > 0:
            self._clear_grace()
        return earned

    def spend(self, amount: float, *, reason: str = "search") -> bool:
        """Always succeeds; the first time the balance reaches <= 0 the 72 h grace clock starts."""
        if amount <= 0:
            raise ValueError(f"amount must be positive, got {amount}")
        with self._lock:
            try:
                self._conn.execute("BEGIN IMMEDIATE")
                now = time.time()
                self._conn.execute("INSERT INTO credit_spending (amount, reason, timestamp) VALUES (?, ?, ?)",
                                   (amount, reason, now))
                if self.balance() <= 0 and self._grace_start() is None:
                    self._conn.execute("UPDATE credit_grace SET grace_start = ? WHERE id = 1", (now,))
                    logger.info("grace_period_started")
                self._conn.execute("COMMIT")
            except Exception:
                with contextlib.suppress(Exception):
                    self._conn.execute("ROLLBACK")
                raise
        return True

    # ------------------------------------------------------------------ balances / state
    def _sum(self, sql: str, args: tuple = ()) -> float:
        return float(self._conn.execute(sql, args).fetchone()[0])

    def total_earned(self) -> float:
        return self._sum("SELECT COALESCE(SUM(credits), 0) FROM credit_entries")

    def total_spent(self) -> float:
        return self._sum("SELECT COALESCE(SUM(amount), 0) FROM credit_spending")

    def balance(self) -> float:
        return self.total_earned() - self.total_spent()

    def debt_amount(self) -> float:
        return max(0.0, -self.balance())

    def _grace_start(self) -> float | None:
        row = self._conn.execute("SELECT grace_start FROM credit_grace WHERE id = 1").fetchone()
        return float(row[0]) if row is not None and row[0] is not None else None

    def _clear_grace(self) -> None:
        if self._grace_start() is not None:
            with self._lock:
                self._conn.execute("UPDATE credit_grace SET grace_start = NULL WHERE id = 1")
                self._conn.commit()

    def credit_state(self, *, now: float | None = None) -> CreditState:
        if self.balance() > 0:
            self._clear_grace()
            return CreditState.NORMAL
        start = self._grace_start()
        if start is None:
            return CreditState.NORMAL
        hours = ((now or time.time()) - start) / 3600.0
        return CreditState.GRACE if hours <= GRACE_PERIOD_HOURS else CreditState.DEBT

    def grace_remaining_hours(self, *, now: float | None = None) -> float | None:
        if self.credit_state(now=now) != CreditState.GRACE:
            return None
        start = self._grace_start()
        return None if start is None else max(0.0, GRACE_PERIOD_HOURS - ((now or time.time()) - start) / 3600.0)

    def search_allowance(self, *, now: float | None = None) -> SearchAllowance:
        state = self.credit_state(now=now)
        cost = self.search_cost() * (DEBT_COST_MULTIPLIER if state == CreditState.DEBT else 1.0)
        return SearchAllowance(state, cost, self.grace_remaining_hours(now=now), self.debt_amount())

    # ------------------------------------------------------------------ score / tier
    def _llm_split(self) -> tuple[float, float]:
        q = ",".join("?" * len(_LLM_VALUES))
        llm = self._sum(f"SELECT COALESCE(SUM(credits), 0) FROM credit_entries WHERE action IN ({q})", _LLM_VALUES)
        non = self._sum(f"SELECT COALESCE(SUM(credits), 0) FROM credit_entries WHERE action NOT IN ({q})", _LLM_VALUES)
        return llm, non

    def contribution_score(self) -> float:
        llm, non = self._llm_split()
        total = llm + non
        if total > 0 and llm / total > LLM_CREDIT_CAP_RATIO:
            llm = non * (LLM_CREDIT_CAP_RATIO / (1 - LLM_CREDIT_CAP_RATIO))
        return non + llm

    def tier(self) -> ContributionTier:
        return _score_to_tier(self.contribution_score())

    def search_cost(self) -> float:
        return _tier_search_cost(self.tier())

    def stats(self) -> LedgerStats:
        llm, non = self._llm_split()
        earned, spent, score = self.total_earned(), self.total_spent(), self.contribution_score()
        allow = self.search_allowance()
        return LedgerStats(round(earned, 4), round(spent, 4), round(earned - spent, 4), round(score, 4),
                           _score_to_tier(score), allow.search_cost, round(llm, 4), round(non, 4), allow.state,
                           allow.grace_remaining_hours, allow.debt_amount, self._owner_email)

    # ------------------------------------------------------------------ listings
    @staticmethod
    def _entry(r) -> CreditEntry:
        return CreditEntry(*r)

    def recent_entries(self, *, limit: int = 50) -> list[CreditEntry]:
        rows = self._conn.execute(f"SELECT {_ENTRY_COLS} FROM credit_entries ORDER BY timestamp DESC LIMIT ?",
                                  (limit,)).fetchall()
        return [self._entry(tuple(r)) for r in rows]

    def signed_entries(self) -> list[CreditEntry]:
        rows = self._conn.execute(f"SELECT {_ENTRY_COLS} FROM credit_entries WHERE entry_hash != '' AND "
                                  "signature != '' ORDER BY timestamp ASC").fetchall()
        return [self._entry(tuple(r)) for r in rows]

    def earnings_by_action(self) -> list[tuple[str, float]]:
        rows = self._conn.execute("SELECT action, COALESCE(SUM(credits), 0) AS total FROM credit_entries "
                                  "WHERE credits > 0 GROUP BY action ORDER BY total DESC").fetchall()
        return [(r[0], r[1]) for r in rows]
