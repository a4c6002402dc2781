"""Credit constants and value types (reference infomesh/credits/types.py:19-149)."""
from __future__ import annotations

from dataclasses import dataclass
from enum import StrEnum


class ActionType(StrEnum):
    CRAWL = "crawl"
    QUERY_PROCESS = "query_process"
    DOC_HOSTING = "doc_hosting"
    NETWORK_UPTIME = "network_uptime"
    LLM_SUMMARIZE_OWN = "llm_own"
    LLM_SUMMARIZE_PEER = "llm_peer"
    GIT_DOCS = "git_docs"
    GIT_FIX = "git_fix"
    GIT_FEATURE = "git_feature"
    GIT_MAJOR = "git_major"


ACTION_WEIGHTS: dict[ActionType, float] = {
    ActionType.CRAWL: 1.0, ActionType.QUERY_PROCESS: 0.5, ActionType.DOC_HOSTING: 0.1,
    ActionType.NETWORK_UPTIME: 0.5, ActionType.LLM_SUMMARIZE_OWN: 1.5, ActionType.LLM_SUMMARIZE_PEER: 2.0,
    ActionType.GIT_DOCS: 1_000.0, ActionType.GIT_FIX: 10_000.0, ActionType.GIT_FEATURE: 50_000.0,
    ActionType.GIT_MAJOR: 100_000.0,
}
GIT_CONTRIBUTION_LEGACY = "git_contrib"
_GIT_ACTIONS = frozenset({ActionType.GIT_DOCS, ActionType.GIT_FIX, ActionType.GIT_FEATURE, ActionType.GIT_MAJOR})
_LLM_ACTIONS = frozenset({ActionType.LLM_SUMMARIZE_OWN, ActionType.LLM_SUMMARIZE_PEER})


class ContributionTier(StrEnum):
    TIER_1 = "tier_1"   # score < 100
    TIER_2 = "tier_2"   # 100 .. 999
    TIER_3 = "tier_3"   # >= 1000


# (minimum score, tier, search cost) — highest first
TIER_THRESHOLDS: list[tuple[float, ContributionTier, float]] = [
    (1000.0, ContributionTier.TIER_3, 0.033), (100.0, ContributionTier.TIER_2, 0.050),
    (0.0, ContributionTier.TIER_1, 0.100)]
LLM_CREDIT_CAP_RATIO: float = 0.60
GRACE_PERIOD_HOURS: float = 72.0
DEBT_COST_MULTIPLIER: float = 2.0


class CreditState(StrEnum):
    NORMAL = "normal"
    GRACE = "grace"
    DEBT = "debt"


@dataclass(frozen=True)
class CreditEntry:
    entry_id: int
    action: str
    quantity: float
    weight: float
    multiplier: float
    credits: float
    timestamp: float
    note: str
    entry_hash: str = ""
    signature: str = ""


@dataclass(frozen=True)
class SearchAllowance:
    """Search is never blocked — only its cost changes."""
    state: CreditState
    search_cost: float
    grace_remaining_hours: float | None
    debt_amount: float


@dataclass(frozen=True)
class LedgerStats:
    total_earned: float
    total_spent: float
    balance: float
    contribution_score: float
    tier: ContributionTier
    search_cost: float
    llm_credits: float
    non_llm_credits: float
    credit_state: CreditState = CreditState.NORMAL
    grace_remaining_hours: float | None = None
    debt_amount: float = 0.0
    owner_email: str = ""
