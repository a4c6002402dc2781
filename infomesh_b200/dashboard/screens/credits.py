"""Credits tab: balance / tier / allowance, earnings by action and the latest ledger entries; posts ``CreditEarned`` when
the total earned grows (reference infomesh/dashboard/screens/credits.py:30-386)."""
from __future__ import annotations

import time

from textual.app import ComposeResult
from textual.containers import Vertical
from textual.message import Message
from textual.widgets import Static

from infomesh_b200.dashboard import utils as U
from infomesh_b200.dashboard.widgets import BarChart


def _read_ledger(config):
    """(stats, allowance, earnings by action, recent entries), or None when there is no ledger yet."""
    path = config.node.data_dir / "credits.db"
    if not path.exists():
        return None
    from infomesh_b200.credits.ledger import CreditLedger

    led = CreditLedger(path)
    try:
        return led.stats(), led.search_allowance(), dict(led.earnings_by_action()), led.recent_entries(limit=8)
    finally:
        led.close()


class _LedgerPanel:
    """Mixin: a panel that can also refresh itself straight from the ledger file when given a config
    (``BalancePanel(config)`` / ``TransactionTable(config)``, the reference's form) -- or be fed by its parent pane."""
    config = None

    def __init__(self, config=None, **kw):
        super().__init__("", **kw)
        self.config = config if config is not None and hasattr(config, "node") else None

    def on_mount(self) -> None:
        if self.config is not None:
            self.refresh_data()

    def refresh_data(self) -> None:
        if self.config is None:
            return
        try:
            data = _read_ledger(self.config)
        except Exception:  # noqa: BLE001
            data = None
        if data is not None:
            self._show_ledger(*data)


class BalancePanel(_LedgerPanel, Static):
    def _show_ledger(self, stats, allowance, by_action, recent) -> None:
        self.show(stats, allowance)

    def show(self, stats, allowance) -> None:
        al = allowance
        extra = (f"  ·  grace {al.grace_remaining_hours:.0f}h left" if al.state.value == "grace" and al.grace_remaining_hours is not None
                 else f"  ·  debt {al.debt_amount:.2f}" if al.state.value == "debt" else "")
        self.update(f"balance [bold green]{stats.balance:,.2f}[/]  ·  {U.tier_label(stats.tier)}  ·  earned {stats.total_earned:,.2f}  "
                    f"spent {stats.total_spent:,.2f}\nsearch cost {al.search_cost:.3f} ({al.state.value}){extra}  ·  "
                    f"contribution score {stats.contribution_score:,.2f}")


class EarningsBreakdownPanel(_LedgerPanel, BarChart):
    def _show_ledger(self, stats, allowance, by_action, recent) -> None:
        self.show(by_action)

    def show(self, by_action: dict[str, float]) -> None:
        self.set_items(sorted(((str(k), float(v)) for k, v in by_action.items()), key=lambda kv: -kv[1])[:8])


class TransactionTable(_LedgerPanel, Static):
    """The most recent ledger entries, newest first."""

    def _show_ledger(self, stats, allowance, by_action, recent) -> None:
        self.show(recent)

    def show(self, entries) -> None:
        if not entries:
            self.update("[dim]no transactions yet[/]")
            return
        rows = [f"[dim]{'time':<9}{'action':<16}{'qty':>6}{'credits':>10}  note[/]"]
        for e in entries:
            note = (e.note or "")[:48].replace("[", "\\[")
            rows.append(f"{time.strftime('%H:%M:%S', time.localtime(e.timestamp)):<9}{e.action:<16}{e.quantity:>6.1f}{e.credits:>+10.2f}  {note}")
        self.update("\n".join(rows))


class CreditsPane(Vertical):
    class CreditEarned(Message):
        def __init__(self, amount: float):
            super().__init__()
            self.amount = amount

    def __init__(self, config, **kw):
        super().__init__(**kw)
        self.config = config
        self._last_earned: float | None = None

    def compose(self) -> ComposeResult:
        yield BalancePanel("", id="cd-head")
        yield Static("[bold]Earnings by action[/]")
        yield EarningsBreakdownPanel("", id="cd-actions")
        yield Static("[bold]Recent transactions[/]")
        yield TransactionTable("", id="cd-tx")

    def on_mount(self) -> None:
        self.refresh_data()
        self.set_interval(3.0, self.refresh_data)

    def refresh_data(self) -> None:
        head = self.query_one(BalancePanel)
        try:
            data = _read_ledger(self.config)
        except Exception as exc:  # noqa: BLE001
            head.update(f"[red]ledger unavailable: {exc}[/]")
            return
        if data is None:
            head.update("[dim]No credit history yet — start crawling to earn credits.[/]")
            return
        s, al, by_action, recent = data
        head.show(s, al)
        self.query_one(EarningsBreakdownPanel).show(by_action)
        self.query_one(TransactionTable).show(recent)
        if self._last_earned is not None and s.total_earned > self._last_earned:
            self.post_message(self.CreditEarned(s.total_earned - self._last_earned))
        self._last_earned = s.total_earned
